"""diagnostic: per-impression scores of the mirror's test path + device vs host ranking metrics"""
import sys, tempfile, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import test_ref_pinned as t
from mnexp_b200 import settings, synth, task, metrics, utils as mu
from sklearn.metrics import roc_auc_score
SH = t.SH
d = tempfile.mkdtemp(); synth.write_dataset(d, SH)
cfg = settings.Config(dict(task='Seq2VecPaperSoftmaxId', arch='igru', score_model='dot', input_training_data_path=d,
    title_shape=SH.L, window_size=SH.W, negative_samples=SH.K, batch_size=SH.B, textual_embedding_dim=SH.E,
    title_filter_shape=(SH.F, SH.k), user_embedding_dim=SH.U, debug=True, dropout=0.0, precision='fp32',
    validation_impression=5, testing_impression=5, epochs=2, training_step=3, validation_step=2, learning_rate=0.001,
    learning_rate_decay=0.2, sparse_user_adam=False))
h = task.get(cfg)
h.build_model(0)
t._load_paper_weights('main-train')(h)
h.model, h.test_model = h.test_model, h.model
preds, trues = [], []
for _, (p, y) in zip(range(5), h.test):
    preds.append(np.asarray(p).reshape(-1)); trues.append(np.asarray(y).reshape(-1))
np.set_printoptions(precision=9, linewidth=200)
dev = metrics.ranking_metrics(preds, trues)
for i, (p, y) in enumerate(zip(preds, trues)):
    host = [roc_auc_score(y, p), mu.ndcg_score(y, p, 10), mu.ndcg_score(y, p, 5), mu.mrr_score(y, p)]
    print(i, 'labels', y.astype(int), 'scores', p)
    print('   device', dev[i], 'host', np.array(host), 'argsort desc', np.argsort(p)[::-1])
