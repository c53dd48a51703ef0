"""Host-side data path of the reference's task/seq2vec.py: Window / Impression / News, DocMeta loading and the
train / valid / test batchers (task/seq2vec.py:16-216).  Integer paths are bit-exact restatements."""
import logging

import numpy as np

from .. import document, settings, utils


def _stack_columns(samples):
    """list of sample tuples -> one stacked numpy array per tuple position."""
    return [np.stack(column) for column in zip(*samples)]


class Seq2Vec:
    """Base task handler: loads DocMeta, exposes `train` / `valid` / `test` batch generators over the subclass's
    `train_gen` / `valid_gen` / `test_gen` sample generators (task/seq2vec.py:16-216)."""

    class Window:
        """The last `window_size` clicked doc ids, oldest first, unfilled slots = doc 0 (left padding);
        `count` = number of pushes so far (task/seq2vec.py:17-53)."""
        __slots__ = ['count', 'docs', 'click_history']

        def __init__(self, docs, window_size):
            self.docs, self.count = docs, 0
            self.click_history = [0] * window_size

        def get_ids(self):
            return np.asarray(self.click_history, dtype=np.int32)

        def get_title(self, elimination=None):
            hidden = elimination or ()
            return np.stack([self.docs[0 if i in hidden else i].title for i in self.click_history])

        def push(self, doc):
            del self.click_history[0]
            self.click_history.append(doc)
            self.count += 1

        @property
        def window_size(self):
            return len(self.click_history)

    class Impression:
        """'p1 p2#TAB#n1 n2 n3' -> clicked / shown-not-clicked doc ids."""
        __slots__ = ['pos', 'neg']

        def __init__(self, d):
            clicked, shown = d.split('#TAB#')[:2]
            self.pos = list(map(int, clicked.split(' ')))
            self.neg = list(map(int, shown.split(' ')))

        def negative_samples(self, n):
            return np.random.choice(self.neg, n)          # with replacement (task/seq2vec.py:61-62)

    class News:
        __slots__ = ['title', 'body']

        def __init__(self, title, body):
            self.title, self.body = title, body

    def _extract_impressions(self, x):
        return [self.Impression(d) for d in x.split('#N#') if not (d.startswith('#TAB#') or d.endswith('#TAB#'))]

    def __init__(self, config: settings.Config):
        self.is_training = True
        self.config = config
        for load in (self._load_docs, self._load_users, self._load_data):
            load()

    def _load_docs(self):
        """DocMeta.tsv (tab-separated: name, id, vertical, subvertical, title tokens, body tokens) ->
        {doc id: News(title (L,), body)} plus the all-zero pad document 0 (task/seq2vec.py:85-110)."""
        logging.info('[+] loading docs metadata')
        to_title = document.DocumentParser(document.parse_document(), document.pad_document(1, self.config.title_shape))
        self.docs = {}
        with open(self.config.doc_meta_input) as file:
            for line in file:
                cols = line.rstrip('\n').split('\t')
                self.docs[int(cols[1])] = self.News(to_title(cols[4])[0], None)
        self.doc_count = 1 + max(self.docs)
        self.docs[0] = self.News(np.zeros(self.config.title_shape), None)
        logging.info('[-] loaded docs metadata')

    def doc_token_table(self):
        """(doc_count, L) int32 table for the device-side token gather (ids absent from DocMeta stay all-zero)."""
        tab = np.zeros((self.doc_count, self.config.title_shape), dtype=np.int32)
        for i, d in self.docs.items():
            tab[i] = d.title.astype(np.int32)
        return tab

    def _load_users(self):
        pass

    def _load_data(self):
        self.training_step = self.config.training_step
        self.validation_step = self.config.validation_step

    @property
    def train(self):
        """Batches drawn from a shuffle pool of 100 * batch_size samples: fill the pool, shuffle it, emit its first
        batch_size samples, keep the rest (task/seq2vec.py:182-193).  Yields ([inputs...], labels)."""
        bs = self.config.batch_size
        pool, samples = [], self.train_gen()
        while True:
            pool.append(next(samples))
            if len(pool) < 100 * bs:
                continue
            np.random.shuffle(pool)
            *inputs, labels = _stack_columns(pool[:bs])
            del pool[:bs]
            yield inputs, labels

    @property
    def valid(self):
        samples = self.valid_gen()
        while True:
            *inputs, labels = _stack_columns([next(samples) for _ in range(self.config.batch_size)])
            yield inputs, labels

    @property
    def test(self):
        """One impression per yield: [predicted scores (n,), labels (n,)] (task/seq2vec.py:202-206)."""
        for impression in self.test_gen():
            *inputs, labels = _stack_columns(impression)
            yield [self.model.predict(inputs).reshape(-1), labels]

    def build_model(self, epoch):
        if epoch == 0:
            self._build_model()
        return self.model

    def callback(self, epoch):
        """task/seq2vec.py:296-322 — inherited by the sigmoid family (Seq2VecPaper / ...Dot / ...Id): learning-rate decay,
        then AUC / nDCG@10 / nDCG@5 / MRR averaged over the first `testing_impression` impressions of `test`; after the
        last epoch once more with is_training = False when a TestData.tsv exists.  The impressions are ranked in one launch
        of the device kernel (mnexp_b200/metrics.py) instead of one sklearn / numpy call each."""
        import os
        from .. import keras_like, metrics
        lr = self.model.optimizer.lr
        keras_like.backend.set_value(lr, keras_like.backend.get_value(lr) * self.config.learning_rate_decay)

        def evaluate():
            preds, trues = [], []
            for _, (y_pred, y_true) in zip(range(self.config.testing_impression), self.test):
                preds.append(np.asarray(y_pred).reshape(-1))
                trues.append(np.asarray(y_true).reshape(-1))
            m = metrics.ranking_metrics(preds, trues)
            rows = [(float(r[0]), float(r[1]), float(r[2]), float(r[3]), np.sum(y), len(y), i)
                    for i, (r, y) in enumerate(zip(m, trues))]
            values = [np.mean(c) for c in zip(*rows)]
            self.last_evaluation = dict(auc=values[0], ndcgx=values[1], ndcgv=values[2], mrr=values[3])
            utils.logging_evaluation(self.last_evaluation)
            utils.logging_evaluation(dict(pos=values[4], size=values[5], num=values[6] * 2 + 1))
        evaluate()
        if epoch == self.config.epochs - 1 and os.path.exists(self.config.testing_data_input):
            self.is_training = False
            evaluate()

    def save_model(self):
        """task/seq2vec.py:324-327 (the paper classes override it with a no-op, task/paper.py:254-255)."""
        logging.info('[+] saving models')
        utils.save_model(self.config.model_output, self.model)
        logging.info('[-] saved models')

    def load_model(self):
        """Restore the weights written by save_model into the built model."""
        utils.load_model(self.config.model_output).apply_to(self.build_model(0) if getattr(self, 'model', None) is None else self.model)
