"""Decomposed inference pipeline — mirror of the reference's task/test_pipeline.py (TestPipeline / TestPipelineProduct).

  test_doc_vec()        doc_encoder over every line `doc \\t title` of pipeline_inputs[0]        (:37-75)
  test_user_vec()       user_encoder over `user_id \\t user_type \\t click#N#click...` lines of
                        pipeline_inputs[1], history vectors taken from the cached doc vectors,
                        unknown documents left at zero                                           (:77-131)
  test_user_doc_score() dot score for every `user_id \\t user_type \\t doc` pair of
                        pipeline_inputs[2], written as `user_id \\t user_type \\t doc \\t score`      (:163-211)
  test_correct()        full model vs pipeline on one (user, doc) pair                          (:214-265)

The arithmetic runs on the GPU through the C-ABI (lstur_encode_docs / lstur_forward_docvecs); this file is host logic.
Differences from the reference, which predates the user-ID models: the model comes from a task handler instead of a
Keras json+pkl pair (`load_model(handler)`), and for the Id models the user index is the integer `user_id` column.
"""
import numpy as np
import torch

from .. import document


class TestPipeline:
    def __init__(self, config):
        self.config = config
        self.handler = None
        self.doc_vec, self.user_vec = {}, {}

    def load_model(self, handler=None):
        if handler is None:
            from .. import task
            handler = task.get(self.config)
            handler.build_model(0)
        self.handler = handler
        # the scoring graph: `test_model` of the softmax classes; the sigmoid family's `model` already is one (task/paper.py:228-256)
        self.model = getattr(handler, 'test_model', None) or handler.model
        self._core = handler._core
        self.score_encoder = None

    # ---- inputs
    def _lines(self, idx):
        with open(self.config.pipeline_inputs[idx]) as f:
            for line in f:
                yield line.strip('\n').split('\t')

    def _test_doc_vec_gen(self):
        for line in self._lines(0):
            yield line[0], line[1]

    def _test_user_vec_gen(self):
        for line in self._lines(1):
            yield line[0], line[1], line[2].split('#N#')

    def _test_user_doc_gen(self):
        for line in self._lines(2):
            yield line[0], line[1], line[2]

    def get_doc_parser(self):
        return document.DocumentParser(document.parse_document(), document.pad_document(1, self.config.title_shape))

    def _engine(self):
        return self._core.engine_infer(1)

    # ---- stages
    def test_doc_vec(self):
        parser = self.get_doc_parser()
        docs, titles = [], []
        for doc, text in self._test_doc_vec_gen():
            docs.append(doc)
            titles.append(np.asarray(parser(text)[0], dtype=np.int32).reshape(-1))
        vecs = self.model.get_layer('doc_encoder').predict(np.stack(titles)) if docs else np.zeros((0, 0), np.float32)
        self.doc_vec = {d: v for d, v in zip(docs, vecs)}
        self._doc_index = {d: i + 1 for i, d in enumerate(docs)}                   # row 0 = unknown document
        D = vecs.shape[1] if docs else 0
        self._table = torch.zeros((len(docs) + 1, D), dtype=torch.float32, device='cuda')
        if docs:
            self._table[1:] = torch.from_numpy(np.ascontiguousarray(vecs)).cuda()

    def _user_batches(self, rows, W, R):
        """rows = [(key, user_index, [doc row ids newest-last])] -> batches of R padded rows."""
        for s in range(0, len(rows), R):
            chunk = rows[s:s + R]
            user = np.zeros(R, dtype=np.int32)
            hist = np.zeros((R, W), dtype=np.int32)
            for i, (_, uidx, ids) in enumerate(chunk):
                user[i] = uidx
                ids = ids[-W:]
                if ids:
                    hist[i, W - len(ids):] = ids                                     # left padded, newest last (:100-108)
            yield chunk, user, hist

    def test_user_vec(self):
        eng = self._engine()
        rows, undoc = [], set()
        for uid, utype, clicks in self._test_user_vec_gen():
            ids = []
            for c in clicks:
                if c in self._doc_index:
                    ids.append(self._doc_index[c])
                else:
                    undoc.add(c)
                    ids.append(0)
            try:
                uidx = int(uid)
            except ValueError:
                uidx = 0
            rows.append((uid + utype, uidx, ids))
        self.user_vec = {}
        for chunk, user, hist in self._user_batches(rows, eng.W, eng.B):
            cand = np.zeros((eng.B, eng.C), dtype=np.int32)
            eng.forward_docvecs(eng.to_device_batch(dict(user=user, hist_doc=hist, cand_doc=cand)), self._table)
            uv = eng.view('user_vec').reshape(eng.B, -1).cpu().numpy()
            for i, (key, _, _) in enumerate(chunk):
                self.user_vec[key] = uv[i].copy()
        print(len(undoc))

    def get_score_encoder(self):
        """(user_vec, doc_vec) -> raw score.  TestPipeline (task/test_pipeline.py:133-150): the model's own scorer without its
        final sigmoid — Dense(1)(concat_dense([u ; d])) of the sigmoid family (task/paper.py:222-226), rebuilt from the
        'concat_dense' / 'socre_dense' weights; TestPipelineProduct (:268-283) and the 'dot' models: u . d.  A few thousand
        pairs per call on the host in float64, like the reference's own last stage."""
        if self.score_encoder is None:
            f8 = lambda a: np.asarray(a, np.float64)
            if self._core.score_model == 'dnn' and not isinstance(self, TestPipelineProduct):
                w = self.model._current()
                sh_w, sh_b, so_w, so_b = f8(w['sh_w']), f8(w['sh_b']), f8(w['so_w']), f8(w['so_b'])
                self.score_encoder = lambda u, d: np.maximum(np.concatenate([f8(u), f8(d)], -1) @ sh_w + sh_b, 0.0) @ so_w + so_b
            else:
                self.score_encoder = lambda u, d: np.sum(f8(u) * f8(d), -1, keepdims=True)
        return self.score_encoder

    def test_user_doc_score(self):
        score = self.get_score_encoder()
        users, docs, uv, dv = [], [], [], []
        with open(self.config.pipeline_output, 'w') as ff:
            def flush():
                if not users:
                    return
                out = score(np.stack(uv), np.stack(dv))
                for (uid, utype), d, o in zip(users, docs, out):
                    ff.write(uid + '\t' + utype + '\t' + d + '\t' + str(float(o[0])) + '\n')
                del users[:], docs[:], uv[:], dv[:]
            for uid, utype, doc in self._test_user_doc_gen():
                key = uid + utype
                if key in self.user_vec and doc in self.doc_vec:
                    users.append((uid, utype)); docs.append(doc)
                    uv.append(self.user_vec[key]); dv.append(self.doc_vec[doc])
                    if len(users) == self.config.batch_size:
                        flush()
            flush()

    def test_correct(self):
        """Full model on the first user / first document vs the pipeline's sigmoid(score): returns (pred, sigm)."""
        parser = self.get_doc_parser()
        doc2title = {d: np.asarray(parser(t)[0], dtype=np.int32).reshape(-1) for d, t in self._test_doc_vec_gen()}
        uid, utype, clicks = next(iter(self._test_user_vec_gen()))
        W, L = self.config.window_size, self.config.title_shape
        clicked = np.zeros((1, W, L), dtype=np.int32)
        n = min(len(clicks), W)
        for i in range(-1, -1 - n, -1):
            if clicks[i] in self.doc_vec:
                clicked[0, i] = doc2title[clicks[i]]
        mydoc = next(iter(doc2title))
        x = [clicked, doc2title[mydoc][None]]
        if self.handler.HAS_USER:
            try:
                uidx = int(uid)
            except ValueError:
                uidx = 0
            x = [np.array([uidx])] + x
        pred = self.model.predict(x)
        out = self.get_score_encoder()(self.user_vec[uid + utype][None], self.doc_vec[mydoc][None])
        sigm = 1.0 / (1.0 + np.exp(-out))
        print(pred)
        print(sigm)
        return np.asarray(pred), sigm


class TestPipelineProduct(TestPipeline):
    pass
