"""Host-side data path of the reference's task/seq2vec.py: Window / Impression / News, DocMeta loading and the
train / valid / test batchers (task/seq2vec.py:16-216).  Integer paths are bit-exact restatements."""
import logging

import numpy as np

from .. import document, settings, utils


class Seq2Vec:
    class Window:
        """Sliding click history of `window_size` doc ids, left-padded with doc 0 (task/seq2vec.py:17-53)."""
        __slots__ = ['count', 'docs', 'click_history']

        def __init__(self, docs, window_size):
            self.count = 0
            self.docs = docs
            self.click_history = [0 for _ in range(window_size)]

        def get_title(self, elimination=None):
            if elimination:
                return np.stack([self.docs[i].title if i not in elimination else self.docs[0].title
                                 for i in self.click_history])
            return np.stack([self.docs[i].title for i in self.click_history])

        def get_ids(self):
            return np.asarray(self.click_history, dtype=np.int32)

        def push(self, doc):
            self.click_history.append(doc)
            self.click_history.pop(0)
            self.count += 1

        @property
        def window_size(self):
            return len(self.click_history)

    class Impression:
        __slots__ = ['pos', 'neg']

        def __init__(self, d):
            d = d.split('#TAB#')
            self.pos = [int(k) for k in d[0].split(' ')]
            self.neg = [int(k) for k in d[1].split(' ')]

        def negative_samples(self, n):
            return np.random.choice(self.neg, n)          # with replacement (task/seq2vec.py:61-62)

    class News:
        __slots__ = ['title', 'body']

        def __init__(self, title, body):
            self.title = title
            self.body = body

    def _extract_impressions(self, x):
        return [self.Impression(d) for d in x.split('#N#') if not d.startswith('#TAB#') and not d.endswith('#TAB#')]

    def __init__(self, config: settings.Config):
        self.is_training = True
        self.config = config
        self._load_docs()
        self._load_users()
        self._load_data()

    def _load_docs(self):
        """DocMeta.tsv -> {doc id: News(title (L,), body)} plus the all-zero pad doc 0 (task/seq2vec.py:85-110)."""
        logging.info('[+] loading docs metadata')
        title_parser = document.DocumentParser(document.parse_document(), document.pad_document(1, self.config.title_shape))
        with open(self.config.doc_meta_input) as file:
            docs = [line.strip('\n').split('\t') for line in file]
        self.docs = {int(line[1]): self.News(title_parser(line[4])[0], None) for line in docs}
        self.doc_count = max(self.docs.keys()) + 1
        doc_example = self.docs[self.doc_count - 1]
        self.docs[0] = self.News(np.zeros_like(doc_example.title), None)
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
        """Shuffle pool of 100*batch_size samples, one batch per yield (task/seq2vec.py:182-193)."""
        pool = []
        size = self.config.batch_size * 100
        gen = self.train_gen()
        while True:
            pool.append(next(gen))
            if len(pool) >= size:
                np.random.shuffle(pool)
                batch = [np.stack(x) for x in zip(*pool[:self.config.batch_size])]
                yield batch[:-1], batch[-1]
                pool = pool[self.config.batch_size:]

    @property
    def valid(self):
        gen = self.valid_gen()
        while True:
            batch = [np.stack(x) for x in zip(*(next(gen) for _ in range(self.config.batch_size)))]
            yield batch[:-1], batch[-1]

    @property
    def test(self):
        for b in self.test_gen():
            batch = [np.stack(x) for x in zip(*b)]
            yield [self.model.predict(batch[:-1]).reshape(-1), batch[-1]]

    def build_model(self, epoch):
        if epoch == 0:
            self._build_model()
        return self.model

    def save_model(self):
        """task/seq2vec.py:324-327 (the paper classes override it with a no-op, task/paper.py:254-255)."""
        logging.info('[+] saving models')
        utils.save_model(self.config.model_output, self.model)
        logging.info('[-] saved models')

    def load_model(self):
        """Restore the weights written by save_model into the built model."""
        utils.load_model(self.config.model_output).apply_to(self.build_model(0) if getattr(self, 'model', None) is None else self.model)
