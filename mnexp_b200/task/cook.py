"""`Cook` task handler of the reference (task/cook.py:4-285) on the CUDA engine.

Data: two .npz files (settings.Config.train_npz_input / test_npz_input) with `idx, idx_mask, ch_title (N,W,L), ch_vert,
ch_subvert (N,W), cd_title (N,5,L), cd_vert, cd_subvert (N,5), cd_label (N,5)`; the test file carries one candidate per
row (`cd_title (n,L), cd_vert, cd_subvert (n,), label, user, impr`).  `train()/valid()/test()` return the same
(features, labels) lists; `build_model(epoch)` returns `train_model` ((1+4)-way softmax CE + Adam) and sets `test_model`
(sigmoid score of one candidate), both with the Keras calls main.py's cook loop makes (fit / evaluate / predict /
predict_on_batch / metrics_names / optimizer.lr / summary).  News vector = [title CNN+attention ‖ Vemb[vert] ‖
Semb[subvert]] without the Dense (task/cook.py:99-113); the user id embedding is multiplied by
Dropout(1 - id_keep)(idx_mask) (:139-142).  Every branch of Cook.get_user_encoder is built (vo, avg, gru, igru, iavg,
iatt, ilstm, agru, ingru, inigru, inagru, atgru, algru, :146-193); unknown names raise the reference's NotImplementedError.
"""
import logging

import numpy as np

from .. import keras_like, metrics, settings, synth
from ..engine import COOK_ARCH, LsturEngine

FEATURES = ['idx', 'idx_mask', 'ch_title', 'ch_vert', 'ch_subvert', 'cd_title', 'cd_vert', 'cd_subvert']
N_USER_ROWS = 25000                                       # keras.layers.Embedding(25000, ...), task/cook.py:137


class CookModel:
    """train_model (train=True) or test_model of Cook._build_model, task/cook.py:214-277."""

    def __init__(self, owner, train):
        self.owner, self.is_train = owner, train
        self.metrics_names = ['loss', 'categorical_accuracy'] if train else ['loss', 'auc_roc']

    @property
    def optimizer(self):
        return self.owner.optimizer

    def _batch(self, x, rows, C, training):
        idx, idx_mask, ch_title, ch_vert, ch_subvert, cd_title, cd_vert, cd_subvert = [np.asarray(a)[rows] for a in x]
        n = len(rows)
        scale = np.asarray(idx_mask, dtype=np.float32).reshape(n)
        keep = self.owner.config.id_keep
        scale2 = None
        two = self.owner.config.arch in ('inigru', 'inagru')    # second id embedding with its own Dropout layer (:169-183)
        if two:
            scale2 = scale.copy()
        if training and keep < 1.0:                       # Dropout(1 - id_keep) on the mask (task/cook.py:141-142)
            scale = scale * (np.random.random(n) < keep).astype(np.float32) / keep
            if two:
                scale2 = scale2 * (np.random.random(n) < keep).astype(np.float32) / keep
        b = dict(user=np.asarray(idx).reshape(n), user_scale=scale, hist_tok=ch_title,
                 cand_tok=np.asarray(cd_title).reshape(n, C, -1))
        if two:
            b['user_scale2'] = scale2
        if self.owner.dv:
            b['hist_vert'] = ch_vert
            b['cand_vert'] = np.asarray(cd_vert).reshape(n, C)
        if self.owner.ds:
            b['hist_subvert'] = ch_subvert
            b['cand_subvert'] = np.asarray(cd_subvert).reshape(n, C)
        return b

    def fit(self, x, y, batch_size=32, epochs=1, initial_epoch=0, shuffle=True, verbose=0, **_):
        assert self.is_train
        y = np.asarray(y[0] if isinstance(y, (list, tuple)) else y, dtype=np.float32)
        n, C = y.shape
        h = keras_like.History()
        for epoch in range(initial_epoch, epochs):
            order = np.random.permutation(n) if shuffle else np.arange(n)
            tot, k = np.zeros(2), 0
            eng = self.owner.engine(batch_size, C, training=True)
            eng.lr = self.optimizer.lr.value
            for s in range(0, n, batch_size):
                rows = order[s:s + batch_size]
                m = len(rows)
                lab = np.zeros((batch_size, C), dtype=np.float32)
                lab[:m] = y[rows]
                if m < batch_size:
                    # Keras trains the ragged last batch too (training.py fit_loop / make_batches).  The plan is per batch
                    # size, so the m samples run as batch_size rows: padding rows repeat the first sample under an all-zero
                    # target (zero loss, zero gradient) and the gradient scale is 1 / m
                    rows = np.concatenate([rows, np.full(batch_size - m, rows[0])])
                b = self._batch(x, rows, C, True)
                b['label'] = lab
                db = eng.to_device_batch(b)
                loss = eng.train_step(db, grad_scale=1.0 / m)
                probs = eng.view('probs').reshape(eng.B, eng.C)[:m]
                tot += [float(loss[0]) * batch_size, float((probs.argmax(1) == db['label'][:m].argmax(1)).float().sum())]
                k += m                              # Keras logs the sample-weighted mean over the batches
            h.epoch.append(epoch)
            for name, v in zip(self.metrics_names, tot / max(1, k)):
                h.history.setdefault(name, []).append(float(v))
        return h

    def predict(self, x, batch_size=None, verbose=0, **_):
        n = len(np.asarray(x[0]))
        C = np.asarray(x[5]).shape[1] if self.is_train else 1
        R = self.owner.predict_rows
        eng = self.owner.engine(R, C, training=False)
        out = np.zeros((n, C), dtype=np.float32)
        for s in range(0, n, R):
            rows = np.arange(s, min(n, s + R))
            pad = np.concatenate([rows, np.full(R - len(rows), rows[-1])])
            db = eng.to_device_batch(self._batch(x, pad, C, False))
            probs = eng.forward(db, training=False)
            res = probs if self.is_train else eng.score_sigmoid()
            out[rows] = res[:len(rows)].cpu().numpy()
        return out

    predict_on_batch = predict

    def evaluate(self, x, y, batch_size=None, verbose=0, **_):
        y = np.asarray(y[0] if isinstance(y, (list, tuple)) else y, dtype=np.float64)
        p = self.predict(x).astype(np.float64)
        if self.is_train:
            q = np.clip(p / p.sum(-1, keepdims=True), 1e-7, 1 - 1e-7)
            return [float((-(y * np.log(q)).sum(-1)).mean()), float((p.argmax(1) == y.argmax(1)).mean())]
        q, yy = np.clip(p.reshape(-1), 1e-7, 1 - 1e-7), y.reshape(-1)     # binary_crossentropy + utils.auc_roc
        return [float(-(yy * np.log(q) + (1 - yy) * np.log(1 - q)).mean()), metrics.auc_roc(q, yy)]

    def get_weights_dict(self):
        return self.owner.current_params()

    def summary(self):
        w = self.owner.current_params()
        print('\n'.join('%-12s %-18s %d' % (k, tuple(np.asarray(v).shape), np.asarray(v).size) for k, v in w.items()))
        print('Total params: %d' % sum(np.asarray(v).size for v in w.values()))


class Cook:
    def __init__(self, config: settings.Config):
        self.config = config
        logging.info('[+] loading training data')
        self.training_data = dict(np.load(self.config.train_npz_input))
        logging.info('[-] loaded training data')
        logging.info('[+] loading testing data')
        self.test_data = dict(np.load(self.config.test_npz_input))
        logging.info('[-] loaded testing data')
        self.predict_rows = 256
        self._train_engine, self._infer = None, {}

    def train(self):
        return [self.training_data[x] for x in FEATURES], [self.training_data[x] for x in ['cd_label']]

    def valid(self):
        k = self.config.validation_step
        return [self.test_data[x][:k] for x in FEATURES], [self.test_data[x][:k] for x in ['label']]

    def test(self):
        return [self.test_data[x] for x in FEATURES], [self.test_data[x] for x in ['user', 'impr', 'idx_mask', 'label']]

    def aggregate_test(self, batch_size=None):
        """The evaluation tail of `main.py cook` (main.py:216-297): score the test set with test_model and average the
        ranking metrics per impression, per user, and per in-vocabulary / out-of-vocabulary user.  Returns the four
        Result objects of mnexp_b200.evaluation.aggregate and logs them like the reference."""
        from .. import evaluation, utils
        feature, (users, imprs, mask, y_true) = self.test()
        y_pred = self.test_model.predict(feature, batch_size=batch_size or self.config.batch_size).reshape(-1)
        res = evaluation.aggregate(users, imprs, np.asarray(mask).reshape(-1), y_true, y_pred)
        evaluation.log_aggregate(res, utils.logging_evaluation)
        return res

    def build_model(self, epoch):
        if epoch == 0:
            self._build_model()
        return self.train_model

    def get_score_model(self, u=None, d=None):
        if self.config.score_model not in ('dot', 'dnn', 'ddot'):
            raise NotImplementedError                      # task/cook.py:210-211
        return self.config.score_model

    def _build_model(self):
        c = self.config
        if c.arch not in COOK_ARCH:
            raise NotImplementedError()                    # task/cook.py:193-194
        if c.news_encoder != 'cnnatt':
            raise NotImplementedError()                    # task/cook.py:96-97
        self.get_score_model()
        word_emb = np.load(c.title_embedding_input + '.npy').astype(np.float32)
        self.W, self.L = self.training_data['ch_title'].shape[1:]
        F, k = c.title_filter_shape
        self.dv = c.vertical_embedding_dim if (c.use_vertical and c.use_vertical_type != 's') else 0
        self.ds = c.subvertical_embedding_dim if (c.use_vertical and c.use_vertical_type != 'v') else 0
        sh = synth.Shape('cook', n_users=N_USER_ROWS, n_news=1, vocab=word_emb.shape[0], L=self.L, W=self.W,
                         K=self.training_data['cd_title'].shape[1] - 1, B=c.batch_size, E=word_emb.shape[1], F=F, k=k,
                         U=c.user_embedding_dim, arch='igru')
        syn = {'vo': 'vo', 'avg': 'niavg', 'gru': 'nigru', 'igru': 'ngru', 'agru': 'pgru', 'ingru': 'igru',
               'inigru': 'iicat'}.get(c.arch, c.arch)      # iavg / iatt / ilstm / inagru / atgru / algru keep their names
        P = synth.make_weights(sh, arch=syn, seed=np.random.randint(1 << 30), word_emb=word_emb, keras_orthogonal=True,
                               score_model=c.score_model, cook=True, dv=self.dv, ds=self.ds)
        if not self.dv:
            P.pop('vert_emb', None)
        if not self.ds:
            P.pop('subvert_emb', None)
        self.params = P
        self.optimizer = keras_like.Adam(c.learning_rate)
        self.train_model = CookModel(self, True)
        self.test_model = CookModel(self, False)
        self.title_encoder = None

    def current_params(self):
        return self._train_engine.get_weights_dict() if self._train_engine is not None else self.params

    def _precision(self):
        p = getattr(self.config, 'precision', 'auto')
        if p != 'auto':
            return p
        from .. import _lib
        ks, E, F = self.params['conv_w'].shape
        return 'fp16_tc' if _lib.load().lstur_tc_supported(self.L, E, F, ks) else 'fp32'

    def engine(self, B, C, training):
        c = self.config
        kw = dict(arch=c.arch, flavour='cook', recurrent_activation=c.recurrent_activation, score_model=c.score_model,
                  precision=self._precision())
        if training or self._train_engine is None:
            e = self._train_engine
            if e is None or ((e.B != B or e.C != C) and training):
                if e is not None:
                    self.params = e.get_weights_dict()
                self._train_engine = LsturEngine(self.params, B, self.W, C, self.L, dropout=c.dropout, lr=c.learning_rate,
                                                 training=True, sparse_user_adam=bool(c.sparse_user_adam),
                                                 trainable_word_emb=bool(c.textual_embedding_trainable), **kw)
                if e is not None:
                    self._train_engine.adopt_state_from(e)      # keep Adam moments / step count / dropout seed counter
                self._infer = {}
            if training:
                return self._train_engine
        key = (B, C)
        if key not in self._infer:
            self._infer[key] = LsturEngine(self.params, B, self.W, C, self.L, dropout=0.0, training=False,
                                           share_weights_from=self._train_engine, **kw)
        return self._infer[key]

    def callback(self, epoch):
        if epoch in self.config.lrd_on_epochs:             # task/cook.py:282-285
            self.optimizer.lr.value = self.optimizer.lr.value * self.config.learning_rate_decay
