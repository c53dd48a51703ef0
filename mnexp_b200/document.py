"""Title token parsing with the call surface of the reference's document.py:12-54.

`DocumentParser(step, step, ...)` composes steps left to right; `parse_document` turns the DocMeta text field into
token-id sentences; `pad_document(size, length)` lays the first `size` non-empty sentences into a zero matrix — this is
the (L,) token layout every title of the hot path has (float64, like the reference's np.zeros buffer)."""
from functools import reduce

import numpy as np


class DocumentParser:
    """Callable pipeline: DocumentParser(f, g)(x) == g(f(x))."""

    def __init__(self, *func):
        self.func = tuple(func)

    def __call__(self, doc):
        return reduce(lambda value, step: step(value), self.func, doc)


def parse_document(sep1='#N#', sep2=' '):
    """'3 7 9#N#4 4' -> [[3, 7, 9], [4, 4]]   (document.py:23-27)"""
    def split(text):
        return [list(map(int, sentence.split(sep2))) for sentence in text.split(sep1)]
    return split


def pad_document(size, length):
    """(size, length) float64 matrix of the first `size` non-empty sentences, each cut to `length` tokens and
    right-padded with 0 (document.py:37-54)."""
    def pad(sentences):
        rows = [s[:length] for s in sentences if s][:size]
        out = np.zeros((size, length))
        for k, row in enumerate(rows):
            out[k, :len(row)] = row
        return out
    return pad
