"""Title token parsers — same behaviour as the reference's document.py:12-54 (token layout consumed by the hot path)."""
import numpy as np


class DocumentParser:
    def __init__(self, *func):
        self.func = func

    def __call__(self, doc):
        for f in self.func:
            doc = f(doc)
        return doc


def parse_document(sep1='#N#', sep2=' '):
    """'3 7 9#N#4 4' -> [[3, 7, 9], [4, 4]]   (document.py:23-27)"""
    def f(doc):
        return [[int(x) for x in d.split(sep2)] for d in doc.split(sep1)]
    return f


def pad_document(size, length):
    """First `size` non-empty sentences, right-zero-padded / truncated to `length`; float64 like np.zeros
    (document.py:37-54)."""
    def f(doc):
        result = np.zeros((size, length))
        i = 0
        for d in doc:
            if d:
                n = min(len(d), length)
                result[i, :n] = d[:n]
                i += 1
                if i == size:
                    break
        return result
    return f
