"""task.get(config) -> eval(config.task)(config), as the reference's task/__init__.py:15-16."""
from .paper import (Seq2VecPaper, Seq2VecPaperDot, Seq2VecPaperId, Seq2VecPaperSoftmax,  # noqa: F401
                    Seq2VecPaperSoftmaxDays, Seq2VecPaperSoftmaxDaysId, Seq2VecPaperSoftmaxDaysIdVert,
                    Seq2VecPaperSoftmaxDaysIdVertAlt, Seq2VecPaperSoftmaxDaysIdVertSup, Seq2VecPaperSoftmaxId)
from .cook import Cook  # noqa: F401
from .seq2vec import Seq2Vec  # noqa: F401


def get(config):
    return eval(config.task)(config)
