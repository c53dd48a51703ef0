"""torch-CPU restatement of the reference LSTUR graph with autograd + Keras Adam.

TEST INFRASTRUCTURE — see ``oracle/__init__.py`` ("parity unpinned").  Second,
independent implementation of the semantics in ``lstur_numpy``;
used (a) in float64 for gradient ground truth, (b) in float32 with all host
threads as the timed "reference-equivalent CPU restatement" baseline
 — the literal Keras/TF-1.x path cannot be imported here.

Reference sites: news encoder task/paper.py:132-160; user encoder :584-633;
scorer :443-464; assembly/loss/optimizer :635-665; custom layers
models.py:20-30, 444-492.
"""
import math

import torch

EPS = 1e-7


def hard_sigmoid(x):
    return torch.clamp(0.2 * x + 0.5, 0.0, 1.0)


def news_encoder(tok, P, use_dense=True, dropout=0.0, training=False, drop_x=None, drop_c=None):
    """task/paper.py:141-160.  tok (N,L) long."""
    X = P['word_emb'][tok]
    if drop_x is not None:
        X = X * drop_x
    elif training and dropout > 0:
        X = torch.nn.functional.dropout(X, dropout, True)
    Wc = P['conv_w']
    k = Wc.shape[0]
    N, L, E = X.shape
    pl = (k - 1) // 2
    Xp = torch.nn.functional.pad(X, (0, 0, pl, k - 1 - pl))
    C = P['conv_b'] + sum(Xp[:, j:j + L] @ Wc[j] for j in range(k))
    C = torch.relu(C)
    C = C * (tok != 0).to(C.dtype).unsqueeze(-1)
    m = (C.detach() != 0).any(-1).to(C.dtype)
    C = C * m.unsqueeze(-1)
    if drop_c is not None:
        C = C * drop_c
    elif training and dropout > 0:
        C = torch.nn.functional.dropout(C, dropout, True)
    a = torch.tanh(C @ P['att_w'].reshape(-1) + P['att_b'].reshape(-1)[0])
    e = torch.exp(a) * m
    w = e / (e.sum(-1, keepdim=True) + EPS)
    p = (C * w.unsqueeze(-1)).sum(1)
    return p @ P['dense_w'] + P['dense_b'] if use_dense else p


def gru_last_state(H, h0, Wx, Wh, b, recurrent_activation='hard_sigmoid'):
    """Keras-2.2 GRU, reset_after=False, masked steps carry state."""
    ra = hard_sigmoid if recurrent_activation == 'hard_sigmoid' else torch.sigmoid
    B, W, D = H.shape
    G = Wh.shape[0]
    gm = (H.detach() != 0).any(-1)
    h = H.new_zeros((B, G)) if h0 is None else h0
    XW = H @ Wx + b
    for t in range(W):
        x = XW[:, t]
        z = ra(x[:, :G] + h @ Wh[:, :G])
        r = ra(x[:, G:2 * G] + h @ Wh[:, G:2 * G])
        hh = torch.tanh(x[:, 2 * G:] + (r * h) @ Wh[:, 2 * G:])
        hn = z * h + (1.0 - z) * hh
        h = torch.where(gm[:, t:t + 1], hn, h)
    return h


def lstm_last_state(H, Wx, Wh, b, recurrent_activation='hard_sigmoid'):
    """Keras-2.2 LSTM (gate order i,f,c,o), masked steps carry (h, c) — task/cook.py:161-163."""
    ra = hard_sigmoid if recurrent_activation == 'hard_sigmoid' else torch.sigmoid
    B, W, D = H.shape
    G = Wh.shape[0]
    gm = (H.detach() != 0).any(-1)
    h, c = H.new_zeros((B, G)), H.new_zeros((B, G))
    XW = H @ Wx + b
    for t in range(W):
        a = XW[:, t] + h @ Wh
        i, f, g, o = ra(a[:, :G]), ra(a[:, G:2 * G]), torch.tanh(a[:, 2 * G:3 * G]), ra(a[:, 3 * G:])
        cn = f * c + i * g
        hn = o * torch.tanh(cn)
        c = torch.where(gm[:, t:t + 1], cn, c)
        h = torch.where(gm[:, t:t + 1], hn, h)
    return h


def masked_attention(X, att_w, att_b):
    """SimpleAttentionMaskSupport()(Masking()(X)) over the steps of X (B,T,D) — models.py:474-489."""
    m = (X.detach() != 0).any(-1).to(X.dtype)
    X = X * m.unsqueeze(-1)
    a = torch.tanh(X @ att_w.reshape(-1) + att_b.reshape(-1)[0])
    e = torch.exp(a) * m
    w = e / (e.sum(-1, keepdim=True) + EPS)
    return (X * w.unsqueeze(-1)).sum(1)


def user_encoder(arch, user, H, P, recurrent_activation='hard_sigmoid', u0_scale=None, u2_scale=None):
    """task/paper.py:584-633; cook branches task/cook.py:146-193."""
    u0 = P['user_emb'][user.reshape(-1)] if 'user_emb' in P and arch not in ('nigru', 'niavg', 'att') else None
    # u2_scale: multiplier of the SECOND id table of cook 'inigru' / 'inagru', which sits behind its own Dropout(1 - id_keep)
    # layer (task/cook.py:169-183: two independent draws); defaults to u0_scale (no dropout: both are the id mask)
    u2_scale = u0_scale if u2_scale is None else u2_scale
    if u0 is not None and u0_scale is not None:
        u0 = u0 * u0_scale
    gru = lambda h0: gru_last_state(H, h0, P['gru_wx'], P['gru_wh'], P['gru_b'], recurrent_activation)
    if arch == 'igru':
        return gru(u0)
    if arch == 'gru':
        return torch.cat([gru(None), u0], -1) @ P['con_w'] + P['con_b']
    if arch == 'iigru':          # task/paper.py:614-619: initial state from table 1, concat with table 2, Dense
        u2 = P['user_emb2'][user.reshape(-1)]
        return torch.cat([gru(u0), u2], -1) @ P['con_w'] + P['con_b']
    if arch in ('ngru', 'hgru', 'dgru'):
        return torch.cat([gru(None), u0], -1)
    if arch == 'iicat':          # Seq2VecPaperId 'iigru' (task/paper.py:338-343) / cook 'inigru' (task/cook.py:169-176, where
        u2 = P['user_emb2'][user.reshape(-1)]       # the id mask multiplies the second embedding too)
        return torch.cat([gru(u0), u2 if u2_scale is None else u2 * u2_scale], -1)
    if arch == 'pgru':
        return gru(None) + u0
    if arch == 'nigru':
        return gru(None)
    if arch == 'vo':
        return u0
    if arch == 'niavg':          # models.GlobalAveragePoolingMaskSupport (models.py:422-441) under Masking()
        gm = (H != 0).any(-1).to(H.dtype)
        return H.sum(-2) / (gm.sum(-1, keepdim=True) + EPS)
    u2 = lambda: P['user_emb2'][user.reshape(-1)] * (1.0 if u2_scale is None else u2_scale)
    if arch == 'iavg':
        gm = (H != 0).any(-1).to(H.dtype)
        return torch.cat([H.sum(-2) / (gm.sum(-1, keepdim=True) + EPS), u0], -1)
    if arch == 'att':
        return masked_attention(H, P['uatt_w'], P['uatt_b'])
    if arch == 'iatt':
        return torch.cat([masked_attention(H, P['uatt_w'], P['uatt_b']), u0], -1)
    if arch == 'ilstm':
        return torch.cat([lstm_last_state(H, P['lstm_wx'], P['lstm_wh'], P['lstm_b'], recurrent_activation), u0], -1)
    if arch == 'inagru':
        return gru(u0) + u2()
    if arch == 'atgru':          # task/cook.py:184-190 as written: 2U one-feature "steps" pooled to one scalar (lstur_numpy.py)
        return masked_attention(torch.cat([gru(None), u0], -1).unsqueeze(-1), P['uatt_w'], P['uatt_b'])
    if arch == 'algru':
        al = P['alpha'].reshape(-1)[0]
        return gru(None) * al + u0 * (1.0 - al)
    raise Exception('Unsupport user model')


def score(u, d, P=None, score_model='dot', flavour='paper'):
    if score_model == 'dot':
        return torch.einsum('bu,bcu->bc', u, d)
    if score_model == 'ddot':
        uh = u @ P['su_w'] + P['su_b']
        dh = d @ P['sd_w'] + P['sd_b']
        if flavour == 'paper':
            uh, dh = torch.tanh(uh), torch.tanh(dh)
        return torch.einsum('bu,bcu->bc', uh, dh)
    if score_model == 'dnn':
        j = torch.cat([u[:, None].expand(-1, d.shape[1], -1), d], -1)
        hid = torch.relu(j @ P['sh_w'] + P['sh_b'])
        return (hid @ P['so_w'] + P['so_b'])[..., 0]
    raise NotImplementedError


def categorical_crossentropy(y, p):
    p = p / p.sum(-1, keepdim=True)
    p = torch.clamp(p, EPS, 1.0 - EPS)
    return (-(y * torch.log(p)).sum(-1)).mean()


def weighted_bce(y, p, gain=1.0, negative_samples=4):
    """Seq2Vec.loss, task/seq2vec.py:213-216 (y, p of the same shape)."""
    K = float(negative_samples)
    return -0.5 * (1 + K) * (y * torch.log(p + 1e-8) * gain + (1 - y) * torch.log(1 - p + 1e-8) / K).mean()


def _doc_vectors(tok, P, vert=None, subvert=None, **kw):
    """paper.py doc encoder, or cook.py's [title ‖ Vemb[vert] ‖ Semb[subvert]] (task/cook.py:99-113)."""
    d = news_encoder(tok, P, use_dense='dense_w' in P, **kw)
    parts = [d]
    if vert is not None and 'vert_emb' in P:
        parts.append(P['vert_emb'][torch.as_tensor(vert).long().reshape(-1)])
    if subvert is not None and 'subvert_emb' in P:
        parts.append(P['subvert_emb'][torch.as_tensor(subvert).long().reshape(-1)])
    return torch.cat(parts, -1) if len(parts) > 1 else d


def forward(P, user, clicked_tok, cand_tok, arch='igru', score_model='dot',
            recurrent_activation='hard_sigmoid', dropout=0.0, training=False, aux=False, hist_vert=None,
            hist_subvert=None, cand_vert=None, cand_subvert=None, u0_scale=None, head='softmax', flavour='paper', u2_scale=None):
    """Seq2VecPaperSoftmaxId._build_model — task/paper.py:635-665."""
    B, W, L = clicked_tok.shape
    C = cand_tok.shape[1]
    dh = _doc_vectors(clicked_tok.reshape(B * W, L), P, hist_vert, hist_subvert, dropout=dropout,
                      training=training).reshape(B, W, -1)
    hm = (clicked_tok != 0).any(-1).to(dh.dtype)
    H = dh * hm.unsqueeze(-1)
    u = user_encoder(arch, user, H, P, recurrent_activation, u0_scale=u0_scale, u2_scale=u2_scale)
    dc = _doc_vectors(cand_tok.reshape(B * C, L), P, cand_vert, cand_subvert, dropout=dropout,
                      training=training).reshape(B, C, -1)
    s = score(u, dc, P, score_model, flavour)
    # head='sigmoid': the sigmoid family (Seq2VecPaper / Dot / Id, task/paper.py:222-262), one candidate per row
    probs = torch.softmax(s, -1) if head == 'softmax' else torch.sigmoid(s)
    if aux:
        return dict(probs=probs, logits=s, user_vec=u, cand_vec=dc, hist_vec=H)
    return probs


def vertical_classifier(P, X):
    """Seq2VecPaperSoftmaxDaysIdVertSup.get_vertical_classifier (task/paper.py:948-952): Dense(hidden_dim, relu) ->
    Dense(len(utils.verticals), softmax), applied TimeDistributed to X (..., D)."""
    return torch.softmax(torch.relu(X @ P['vs_w1'] + P['vs_b1']) @ P['vs_w2'] + P['vs_b2'], -1)


def loss_fn(P, user, clicked_tok, cand_tok, label=None, vert_labels=None, aux_gain=1.0, parts=False, **kw):
    """categorical cross-entropy of the click head; with vert_labels = (hist (B,W), cand (B,C)) integer vertical ids also
    the auxiliary head of ...VertSup (task/paper.py:954-990): the classifier runs over [clicked_vec (history-masked) ;
    candidate vectors] and the compiled loss is 1 * CE_click + gain * CE_vertical, the latter a mean over all B*(W+C)
    positions."""
    if vert_labels is None:
        probs = forward(P, user, clicked_tok, cand_tok, **kw)
    else:
        out = forward(P, user, clicked_tok, cand_tok, aux=True, **kw)
        probs = out['probs']
    if label is None:
        label = torch.zeros_like(probs)
        label[:, 0] = 1.0
    main = categorical_crossentropy(label, probs)
    if vert_labels is None:
        return main
    X = torch.cat([out['hist_vec'], out['cand_vec']], 1)                  # (B, W+C, D)
    y = torch.cat([torch.as_tensor(vert_labels[0]).long(), torch.as_tensor(vert_labels[1]).long()], 1)
    vp = vertical_classifier(P, X)
    onehot = torch.nn.functional.one_hot(y, vp.shape[-1]).to(vp.dtype)
    aux = categorical_crossentropy(onehot, vp)
    if parts:
        return main, aux, vp
    return main + aux_gain * aux


def title_cls_loss(P, tok, labels, dropout=0.0, training=False, drop_x=None, drop_c=None, parts=False):
    """vert_model of Seq2VecPaperSoftmaxDaysIdVertAlt (task/paper.py:1128-1136): Dense(n_vert, softmax)(doc_encoder(title)),
    categorical cross-entropy against the one-hot vertical."""
    d = news_encoder(torch.as_tensor(tok).long(), P, use_dense='dense_w' in P, dropout=dropout, training=training,
                     drop_x=drop_x, drop_c=drop_c)
    vp = torch.softmax(d @ P['vcls_w'] + P['vcls_b'], -1)
    onehot = torch.nn.functional.one_hot(torch.as_tensor(labels).long(), vp.shape[-1]).to(vp.dtype)
    loss = categorical_crossentropy(onehot, vp)
    return (loss, vp) if parts else loss


class KerasAdam:
    """keras.optimizers.Adam 2.2.x, dense for every tensor incl. embedding tables."""

    def __init__(self, params, lr=1e-3, b1=0.9, b2=0.999, eps=1e-7):
        self.params, self.lr, self.b1, self.b2, self.eps = params, lr, b1, b2, eps
        self.t = 0
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}

    @torch.no_grad()
    def step(self, grads):
        self.t += 1
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        for k, g in grads.items():
            if g is None:
                continue
            self.m[k].mul_(self.b1).add_(g, alpha=1.0 - self.b1)
            self.v[k].mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
            self.params[k].sub_(lr_t * self.m[k] / (self.v[k].sqrt() + self.eps))


class LsturOracle:
    """Trainable oracle model: params dict of leaf tensors + KerasAdam."""

    def __init__(self, params_np, arch='igru', score_model='dot', dtype=torch.float64, lr=1e-3,
                 trainable_word_emb=False, dropout=0.0, recurrent_activation='hard_sigmoid'):
        self.arch, self.score_model, self.dropout = arch, score_model, dropout
        self.ra = recurrent_activation
        self.P = {k: torch.tensor(v, dtype=dtype) for k, v in params_np.items()}
        self.trainable = [k for k in self.P if k != 'word_emb' or trainable_word_emb]
        for k in self.trainable:
            self.P[k].requires_grad_(True)
        self.opt = KerasAdam({k: self.P[k] for k in self.trainable}, lr=lr)

    def _ints(self, user, clicked_tok, cand_tok):
        t = lambda x: torch.as_tensor(x).long()
        return t(user), t(clicked_tok), t(cand_tok)

    def forward(self, user, clicked_tok, cand_tok, training=False, aux=False):
        u, c, d = self._ints(user, clicked_tok, cand_tok)
        return forward(self.P, u, c, d, arch=self.arch, score_model=self.score_model, recurrent_activation=self.ra,
                       dropout=self.dropout, training=training, aux=aux)

    def loss_and_grads(self, user, clicked_tok, cand_tok, training=False):
        u, c, d = self._ints(user, clicked_tok, cand_tok)
        loss = loss_fn(self.P, u, c, d, arch=self.arch, score_model=self.score_model, recurrent_activation=self.ra,
                       dropout=self.dropout, training=training)
        gs = torch.autograd.grad(loss, [self.P[k] for k in self.trainable], allow_unused=True)
        return loss.detach(), dict(zip(self.trainable, gs))

    def train_step(self, user, clicked_tok, cand_tok, training=True):
        loss, grads = self.loss_and_grads(user, clicked_tok, cand_tok, training=training)
        self.opt.step(grads)
        return float(loss)
