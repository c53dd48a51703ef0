"""ctypes binding of include/lstur_b200.h (the C-ABI boundary).

The prototypes are parsed from the header itself so the binding cannot drift
from the declared ABI.  There is NO fallback: if the shared library is missing
or a symbol is absent this module raises.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, '..', 'include', 'lstur_b200.h')
LIB_PATH = os.path.join(HERE, 'lib', 'liblstur_b200.so')

_SCALARS = {
    'int': ctypes.c_int, 'unsigned': ctypes.c_uint, 'float': ctypes.c_float, 'long long': ctypes.c_longlong,
    'size_t': ctypes.c_size_t, 'cudaStream_t': ctypes.c_void_p, 'void': None,
    'unsigned long long': ctypes.c_ulonglong,
}


class lstur_config(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in
                ('B', 'W', 'C', 'L', 'E', 'F', 'KS', 'use_dense', 'Dd', 'dv', 'ds', 'G', 'Ue', 'U',
                 'arch', 'score_model', 'rec_act', 'precision', 'V', 'n_users', 'n_docs')] + \
               [('dropout', ctypes.c_float), ('save_for_backward', ctypes.c_int), ('n_vert', ctypes.c_int),
                ('n_subvert', ctypes.c_int), ('Hs', ctypes.c_int), ('loss_model', ctypes.c_int),
                ('bce_neg', ctypes.c_int), ('gain', ctypes.c_float), ('trainable_word_emb', ctypes.c_int),
                ('aux_nv', ctypes.c_int), ('aux_hidden', ctypes.c_int), ('aux_gain', ctypes.c_float),
                ('cls_nv', ctypes.c_int)]


class lstur_weights(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ('dense', 'word_emb', 'user_emb', 'doc_tokens', 'doc_vert', 'doc_subvert')]


class lstur_batch(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in
                ('user', 'hist_doc', 'cand_doc', 'hist_tok', 'cand_tok', 'label', 'user_scale', 'hist_vert',
                 'hist_subvert', 'cand_vert', 'cand_subvert', 'user_scale2')]


def parse_header(path=HEADER):
    """-> {name: (restype, [argtypes])} for every function prototype in the header."""
    src = open(path).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    src = re.sub(r'//[^\n]*', '', src)
    src = re.sub(r'typedef struct \w+ \{.*?\} \w+;', '', src, flags=re.S)
    protos = {}
    for m in re.finditer(r'\b(const char\*|unsigned long long|int|void|size_t|long long)\s+(lstur_\w+)\s*\(([^)]*)\)\s*;', src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        argtypes = []
        if args and args != 'void':
            for a in args.split(','):
                a = a.strip()
                if '*' in a:
                    argtypes.append(ctypes.c_void_p)
                    continue
                a = re.sub(r'\bconst\b', '', a).strip()
                ty = ' '.join(a.split()[:-1])
                argtypes.append(_SCALARS[ty])
        restype = ctypes.c_char_p if ret == 'const char*' else _SCALARS[ret]
        protos[name] = (restype, argtypes)
    return protos


class LsturError(RuntimeError):
    pass


_lib = None


def load():
    """Load liblstur_b200.so and bind every prototype of the header (fails loudly)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LsturError('%s not found: run `python -c "import __graft_entry__ as g; g.build()"` '
                         '(there is no CPU fallback)' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in parse_header().items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise LsturError('lstur error %d: %s' % (rc, load().lstur_last_error().decode()))


def call(name, *args):
    """Call an int-returning ABI function and raise on a non-zero status."""
    check(getattr(load(), name)(*args))
