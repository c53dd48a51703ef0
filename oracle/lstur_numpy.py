"""float64 numpy restatement of the reference LSTUR forward path (ground truth).

TEST INFRASTRUCTURE — see ``oracle/__init__.py`` ("parity unpinned").

Every function cites the reference graph-construction site it restates
(paths relative to /root/reference) and the SURVEY.md §9 item that spells out
the Keras-2.2/TF-1.x library semantics ([K]) which are not vendored there.

Parameter dictionary (names used by oracle, engine and tests alike):
  word_emb (V,E)        Embedding(weights=[title_embedding])   task/paper.py:132-138
  conv_w (k,E,F) conv_b (F)   Conv1D(F,k,'same',relu)          task/paper.py:146
  att_w (F,) att_b ()         SimpleAttentionMaskSupport       models.py:456-468
  dense_w (F,U) dense_b (U)   Dense(user_embedding_dim)        task/paper.py:159   (absent in cook.py)
  vert_emb (16,dv) subvert_emb (307,ds)                        task/cook.py:99-103
  user_emb (n_users,Ue)       Embedding(len(self.data),U)      task/paper.py:589
  gru_wx (D,3G) gru_wh (G,3G) gru_b (3G)   keras GRU, gate order z,r,h [K]
  con_w (G+Ue,U) con_b (U)    Dense after concat, arch 'gru'   task/paper.py:598-599
  lstm_wx (D,4G) lstm_wh (G,4G) lstm_b (4G)   keras LSTM, gate order i,f,c,o [K]   task/cook.py:161-163
  uatt_w (Da,) uatt_b ()      SimpleAttentionMaskSupport over the history / over [GRU ; id]   task/cook.py:158-160,184-190
  alpha ()                    models.AlphaAdd                  models.py:540-554, task/cook.py:191-193
  su_w,su_b,sd_w,sd_b         'ddot' scorer Dense layers       task/paper.py:452-455, task/cook.py:206-209
  sh_w,sh_b,so_w,so_b         'dnn' scorer                     task/paper.py:448-451
"""
import numpy as np

EPS = 1e-7  # keras.backend.epsilon() [K]


def hard_sigmoid(x):
    """Keras-2.2 hard_sigmoid: clip(0.2x+0.5,0,1) [K]."""
    return np.clip(0.2 * x + 0.5, 0.0, 1.0)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def token_gather(doc_tokens, doc_ids):
    """Window.get_title: np.stack([docs[i].title ...]) — task/seq2vec.py:25-30."""
    return np.asarray(doc_tokens)[np.asarray(doc_ids)]


# --------------------------------------------------------------------------
# news encoder — task/paper.py:132-160 (cook twin task/cook.py:42-61)
# --------------------------------------------------------------------------
def conv1d_same_relu(X, Wc, bc):
    """Conv1D(F,k,padding='same',strides=1,activation='relu') [K].

    X (N,L,E); Wc (k,E,F); zero pad floor((k-1)/2) left, rest right."""
    N, L, E = X.shape
    k = Wc.shape[0]
    pl = (k - 1) // 2
    Xp = np.zeros((N, L + k - 1, E), dtype=X.dtype)
    Xp[:, pl:pl + L] = X
    C = np.zeros((N, L, Wc.shape[2]), dtype=X.dtype)
    for j in range(k):
        C += Xp[:, j:j + L] @ Wc[j]
    return np.maximum(C + bc, 0.0)


def attention_pool(C, m, att_w, att_b):
    """models.SimpleAttentionMaskSupport.call — models.py:474-489."""
    a = np.tanh(C @ att_w + att_b)                # (N,L)
    e = np.exp(a) * m
    w = e / (e.sum(-1, keepdims=True) + EPS)
    return (C * w[..., None]).sum(1), dict(a=a, w=w)


def news_encoder(tok, P, use_dense=True, drop_x=None, drop_c=None, aux=False):
    """tok (N,L) int -> doc vectors.  drop_x/drop_c: optional explicit inverted
    dropout multipliers (already scaled by 1/(1-p)) for the two Dropout layers
    (task/paper.py:147,158) so a GPU dropout mask can be replayed exactly."""
    tok = np.asarray(tok).astype(np.int64)
    X = P['word_emb'].astype(np.float64)[tok]                   # mask_zero=False: row 0 is a normal row
    if drop_x is not None:
        X = X * drop_x
    C = conv1d_same_relu(X, P['conv_w'].astype(np.float64), P['conv_b'].astype(np.float64))
    C = C * (tok != 0)[..., None]                               # Lambda, task/paper.py:150-155
    m = (C != 0).any(-1).astype(np.float64)                     # Masking() [K]
    C = C * m[..., None]
    if drop_c is not None:
        C = C * drop_c                                          # Dropout keeps the mask [K]
    p, att = attention_pool(C, m, P['att_w'].astype(np.float64).reshape(-1), float(np.asarray(P['att_b']).reshape(-1)[0]))
    d = p @ P['dense_w'].astype(np.float64) + P['dense_b'].astype(np.float64) if use_dense else p
    if aux:
        return d, dict(C=C, m=m, p=p, **att)
    return d


def news_encoder_vert(tok, vert, subvert, P, use_vertical_type='vs', **kw):
    """Cook.get_doc_encoder concat — task/cook.py:99-113: [title ‖ Vemb[vert] ‖ Semb[subvert]]."""
    d = news_encoder(tok, P, use_dense=False, **kw)
    parts = [d]
    if use_vertical_type in ('v', 'vs'):
        parts.append(P['vert_emb'].astype(np.float64)[np.asarray(vert).astype(np.int64)])
    if use_vertical_type in ('s', 'vs'):
        parts.append(P['subvert_emb'].astype(np.float64)[np.asarray(subvert).astype(np.int64)])
    return np.concatenate(parts, -1)


# --------------------------------------------------------------------------
# history mask + GRU — models.py:20-30, task/paper.py:644-645, 584-633
# --------------------------------------------------------------------------
def history_mask(clicked_tok):
    """models.ComputeMasking(0): any(tokens != 0, axis=-1) — models.py:25-27."""
    return (np.asarray(clicked_tok) != 0).any(-1).astype(np.float64)


def gru_last_state(H, h0, Wx, Wh, b, recurrent_activation='hard_sigmoid', aux=False):
    """keras.layers.GRU(G)(Masking()(H), initial_state=h0), return_sequences=False.

    Keras 2.2.x defaults [K]: reset_after=False (reset gate is
    applied BEFORE the recurrent matmul), gate order z,r,h, masked steps carry
    the state.  Masking() recomputes gm = any(H != 0, -1) (task/paper.py:592)."""
    ra = hard_sigmoid if recurrent_activation == 'hard_sigmoid' else sigmoid
    B, W, D = H.shape
    G = Wh.shape[0]
    gm = (H != 0).any(-1)
    h = np.zeros((B, G)) if h0 is None else h0.astype(np.float64).copy()
    steps = []
    for t in range(W):
        x = H[:, t] @ Wx + b
        z = ra(x[:, :G] + h @ Wh[:, :G])
        r = ra(x[:, G:2 * G] + h @ Wh[:, G:2 * G])
        hh = np.tanh(x[:, 2 * G:] + (r * h) @ Wh[:, 2 * G:])
        hn = z * h + (1.0 - z) * hh
        hprev = h
        h = np.where(gm[:, t:t + 1], hn, h)
        steps.append(dict(z=z, r=r, hh=hh, hprev=hprev))
    if aux:
        return h, dict(gm=gm.astype(np.float64), steps=steps)
    return h


def lstm_last_state(H, Wx, Wh, b, recurrent_activation='hard_sigmoid'):
    """keras.layers.LSTM(G)(Masking()(H)) — task/cook.py:161-163.  Keras 2.2.x defaults [K]: activation tanh,
    recurrent_activation hard_sigmoid, gate order i,f,c,o, zero initial (h, c), masked steps carry both."""
    ra = hard_sigmoid if recurrent_activation == 'hard_sigmoid' else sigmoid
    B, W, D = H.shape
    G = Wh.shape[0]
    gm = (H != 0).any(-1)
    h, c = np.zeros((B, G)), np.zeros((B, G))
    for t in range(W):
        a = H[:, t] @ Wx + b + h @ Wh
        i, f, g, o = ra(a[:, :G]), ra(a[:, G:2 * G]), np.tanh(a[:, 2 * G:3 * G]), ra(a[:, 3 * G:])
        cn = f * c + i * g
        hn = o * np.tanh(cn)
        c = np.where(gm[:, t:t + 1], cn, c)
        h = np.where(gm[:, t:t + 1], hn, h)
    return h


def masked_attention(X, att_w, att_b):
    """SimpleAttentionMaskSupport()(Masking()(X)) over the steps of X (B,T,D) — models.py:474-489."""
    m = (X != 0).any(-1).astype(np.float64)
    return attention_pool(X * m[..., None], m, att_w, att_b)[0]


ARCH_INI = ('igru', 'ingru')          # paper.py igru == cook.py ingru


def user_encoder(arch, user, H, P, recurrent_activation='hard_sigmoid', u0_scale=None, u2_scale=None):
    """Seq2VecPaperSoftmaxId.get_user_encoder — task/paper.py:584-633.

    H (B,W,D) already multiplied by the history mask.  u0_scale: optional (B,1)
    multiplier on the user vector (dgru whole-vector dropout :609, cook id_keep
    task/cook.py:141-142)."""
    f8 = lambda k: P[k].astype(np.float64)
    user = np.asarray(user).astype(np.int64).reshape(-1)
    u0 = f8('user_emb')[user] if 'user_emb' in P and arch not in ('nigru', 'niavg', 'att') else None
    u2_scale = u0_scale if u2_scale is None else u2_scale      # second id table: its own Dropout draw (task/cook.py:169-183)
    if u0 is not None and u0_scale is not None:
        u0 = u0 * u0_scale
    gru = lambda h0: gru_last_state(H.astype(np.float64), h0, f8('gru_wx'), f8('gru_wh'), f8('gru_b'), recurrent_activation)
    if arch == 'igru':                       # LSTUR-ini, :612-613
        return gru(u0)
    if arch == 'gru':                        # LSTUR-con + Dense, :596-599
        return np.concatenate([gru(None), u0], -1) @ f8('con_w') + f8('con_b')
    if arch == 'iigru':                      # :614-619
        return np.concatenate([gru(u0), f8('user_emb2')[user]], -1) @ f8('con_w') + f8('con_b')
    if arch in ('ngru', 'hgru', 'dgru'):     # LSTUR-con plain concat, :600-611
        return np.concatenate([gru(None), u0], -1)
    if arch == 'iicat':                      # Seq2VecPaperId 'iigru', task/paper.py:338-343 (cook 'inigru', task/cook.py:169-176)
        return np.concatenate([gru(u0), f8('user_emb2')[user] * (1.0 if u2_scale is None else u2_scale)], -1)
    if arch == 'pgru':                       # :622-624
        return gru(None) + u0
    if arch == 'nigru':                      # :625-626
        return gru(None)
    if arch == 'vo':                         # :620-621
        return u0
    if arch == 'niavg':                      # :627-628, models.GlobalAveragePoolingMaskSupport (models.py:433-435)
        H8 = H.astype(np.float64)
        gm = (H8 != 0).any(-1).astype(np.float64)
        return H8.sum(-2) / (gm.sum(-1, keepdims=True) + 1e-7)
    # ---- cook.py branches (task/cook.py:155-193) and Seq2VecPaper 'att' (task/paper.py:206-208)
    H8 = H.astype(np.float64)
    u2 = lambda: f8('user_emb2')[user] * (1.0 if u2_scale is None else u2_scale)   # the id mask multiplies both tables
    if arch == 'iavg':
        gm = (H8 != 0).any(-1).astype(np.float64)
        return np.concatenate([H8.sum(-2) / (gm.sum(-1, keepdims=True) + 1e-7), u0], -1)
    if arch == 'att':
        return masked_attention(H8, f8('uatt_w').reshape(-1), float(np.asarray(P['uatt_b']).reshape(-1)[0]))
    if arch == 'iatt':
        return np.concatenate([masked_attention(H8, f8('uatt_w').reshape(-1), float(np.asarray(P['uatt_b']).reshape(-1)[0])), u0], -1)
    if arch == 'ilstm':
        return np.concatenate([lstm_last_state(H8, f8('lstm_wx'), f8('lstm_wh'), f8('lstm_b'), recurrent_activation), u0], -1)
    if arch == 'inagru':
        return gru(u0) + u2()
    if arch == 'atgru':
        # task/cook.py:184-190 as WRITTEN: expand_dims(x, -1) of both vectors, concatenated on axis -2, is a sequence of
        # 2U "steps" of ONE feature; Masking drops the zero entries and SimpleAttentionMaskSupport (kernel (1, 1)) pools the
        # 2U scalars into a single number per user — established by running the reference's own code
        # (tests/golden/make_ref_golden.py), not the two-step (B, 2, U) pooling a reader expects
        seq = np.concatenate([gru(None), u0], -1)[..., None]
        return masked_attention(seq, f8('uatt_w').reshape(-1), float(np.asarray(P['uatt_b']).reshape(-1)[0]))
    if arch == 'algru':
        al = float(np.asarray(P['alpha']).reshape(-1)[0])
        return gru(None) * al + u0 * (1.0 - al)
    raise Exception('Unsupport user model')  # task/paper.py:630


# --------------------------------------------------------------------------
# score / loss — task/paper.py:443-464, 655-665; keras losses [K]
# --------------------------------------------------------------------------
def score(u, d, P=None, score_model='dot', flavour='paper'):
    """u (B,Du), d (B,C,Dd) -> logits (B,C)."""
    if score_model == 'dot':
        return np.einsum('bu,bcu->bc', u, d)
    f8 = lambda k: P[k].astype(np.float64)
    if score_model == 'ddot':
        uh = u @ f8('su_w') + f8('su_b')
        dh = d @ f8('sd_w') + f8('sd_b')
        if flavour == 'paper':               # tanh in paper.py:453-454, linear in cook.py:206-207
            uh, dh = np.tanh(uh), np.tanh(dh)
        return np.einsum('bu,bcu->bc', uh, dh)
    if score_model == 'dnn':
        j = np.concatenate([np.broadcast_to(u[:, None], d.shape[:2] + u.shape[-1:]), d], -1)
        hid = np.maximum(j @ f8('sh_w') + f8('sh_b'), 0.0)
        return (hid @ f8('so_w') + f8('so_b'))[..., 0]
    raise NotImplementedError


def softmax(s):
    e = np.exp(s - s.max(-1, keepdims=True))
    return e / e.sum(-1, keepdims=True)


def categorical_crossentropy(y, p):
    """keras.losses.categorical_crossentropy on probabilities [K]."""
    p = p / p.sum(-1, keepdims=True)
    p = np.clip(p, EPS, 1.0 - EPS)
    return float((-(y * np.log(p)).sum(-1)).mean())


def weighted_bce(y, p, gain, K):
    """Seq2Vec.loss — task/seq2vec.py:213-216."""
    return float(-0.5 * (1 + K) * np.mean(y * np.log(p + 1e-8) * gain + (1 - y) * np.log(1 - p + 1e-8) / K))


def _doc_vectors(tok, P, vert=None, subvert=None):
    """paper.py doc encoder, or cook.py's [title ‖ Vemb[vert] ‖ Semb[subvert]] when vertical ids are given."""
    d = news_encoder(tok, P, use_dense='dense_w' in P)
    parts = [d]
    if vert is not None and 'vert_emb' in P:
        parts.append(P['vert_emb'].astype(np.float64)[np.asarray(vert).astype(np.int64).reshape(-1)])
    if subvert is not None and 'subvert_emb' in P:
        parts.append(P['subvert_emb'].astype(np.float64)[np.asarray(subvert).astype(np.int64).reshape(-1)])
    return np.concatenate(parts, -1) if len(parts) > 1 else d


def lstur_forward(P, user, clicked_tok, cand_tok, arch='igru', score_model='dot',
                  recurrent_activation='hard_sigmoid', aux=False, hist_vert=None, hist_subvert=None, cand_vert=None,
                  cand_subvert=None, u0_scale=None, flavour='paper', u2_scale=None):
    """Seq2VecPaperSoftmaxId._build_model forward — task/paper.py:635-665.

    clicked_tok (B,W,L), cand_tok (B,C,L) -> softmax probs (B,C) and the
    test-model sigmoid scores."""
    B, W, L = clicked_tok.shape
    C = cand_tok.shape[1]
    dh = _doc_vectors(clicked_tok.reshape(B * W, L), P, hist_vert, hist_subvert).reshape(B, W, -1)
    hm = history_mask(clicked_tok)
    H = dh * hm[..., None]                                    # task/paper.py:644-645, task/cook.py:250
    u = user_encoder(arch, user, H, P, recurrent_activation, u0_scale=u0_scale, u2_scale=u2_scale)
    dc = _doc_vectors(cand_tok.reshape(B * C, L), P, cand_vert, cand_subvert).reshape(B, C, -1)
    s = score(u, dc, P, score_model, flavour)
    out = dict(probs=softmax(s), logits=s, sigmoid=sigmoid(s), user_vec=u, cand_vec=dc, hist_vec=H, hist_mask=hm)
    return out if aux else out['probs']


def lstur_loss(P, user, clicked_tok, cand_tok, label=None, **kw):
    probs = lstur_forward(P, user, clicked_tok, cand_tok, **kw)
    if label is None:
        label = np.zeros_like(probs)
        label[:, 0] = 1.0                                     # positive first, task/paper.py:529
    return categorical_crossentropy(label, probs)


# --------------------------------------------------------------------------
# Adam — keras.optimizers.Adam (2.2.x) [K]
# --------------------------------------------------------------------------
def adam_step(p, g, m, v, t, lr, b1=0.9, b2=0.999, eps=1e-7):
    """One dense Keras-2.2 Adam update; t is the 1-based step index after increment."""
    lr_t = lr * np.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    m = b1 * m + (1.0 - b1) * g
    v = b2 * v + (1.0 - b2) * g * g
    p = p - lr_t * m / (np.sqrt(v) + eps)
    return p, m, v


# --------------------------------------------------------------------------
# pure-Python loop versions for hand-checkable micro cases
# --------------------------------------------------------------------------
def news_encoder_loops(tok, P):
    """Scalar-loop restatement of news_encoder (use only for tiny shapes)."""
    import math
    tok = np.asarray(tok).astype(int)
    N, L = tok.shape
    We, Wc, bc = P['word_emb'], P['conv_w'], P['conv_b']
    k, E, F = Wc.shape
    ka, ba = np.asarray(P['att_w']).reshape(-1), float(np.asarray(P['att_b']).reshape(-1)[0])
    Wd, bd = P['dense_w'], P['dense_b']
    U = Wd.shape[1]
    out = np.zeros((N, U))
    for n in range(N):
        C = [[0.0] * F for _ in range(L)]
        for t in range(L):
            for f in range(F):
                acc = float(bc[f])
                for j in range(k):
                    s = t + j - (k - 1) // 2
                    if 0 <= s < L:
                        for e in range(E):
                            acc += float(We[tok[n, s], e]) * float(Wc[j, e, f])
                acc = max(acc, 0.0)
                if tok[n, t] == 0:
                    acc = 0.0
                C[t][f] = acc
        m = [1.0 if any(c != 0.0 for c in C[t]) else 0.0 for t in range(L)]
        e_ = [math.exp(math.tanh(sum(C[t][f] * float(ka[f]) for f in range(F)) + ba)) * m[t] for t in range(L)]
        S = sum(e_) + EPS
        p = [sum(e_[t] / S * C[t][f] * m[t] for t in range(L)) for f in range(F)]
        for u in range(U):
            out[n, u] = float(bd[u]) + sum(p[f] * float(Wd[f, u]) for f in range(F))
    return out


def gru_loops(H, h0, Wx, Wh, b):
    """Scalar-loop Keras-2.2 GRU (hard_sigmoid, reset-before, masked carry)."""
    import math
    B, W, D = H.shape
    G = Wh.shape[0]
    hs = lambda x: min(1.0, max(0.0, 0.2 * x + 0.5))
    out = np.zeros((B, G))
    for bi in range(B):
        h = [0.0] * G if h0 is None else [float(x) for x in h0[bi]]
        for t in range(W):
            if not any(float(x) != 0.0 for x in H[bi, t]):
                continue
            xp = [float(b[c]) + sum(float(H[bi, t, d]) * float(Wx[d, c]) for d in range(D)) for c in range(3 * G)]
            z = [hs(xp[j] + sum(h[k] * float(Wh[k, j]) for k in range(G))) for j in range(G)]
            r = [hs(xp[G + j] + sum(h[k] * float(Wh[k, G + j]) for k in range(G))) for j in range(G)]
            hh = [math.tanh(xp[2 * G + j] + sum(r[k] * h[k] * float(Wh[k, 2 * G + j]) for k in range(G))) for j in range(G)]
            h = [z[j] * h[j] + (1.0 - z[j]) * hh[j] for j in range(G)]
        out[bi] = h
    return out
