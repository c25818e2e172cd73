"""`.pr` result files: a pickled dict {name: object} (boundary, BasicFunctionsSJR.py:25-89 of the reference)."""
import os
import pickle


def mkdir(path):
    if path and not os.path.isdir(path):
        os.makedirs(path, exist_ok=True)


def save_pr(path, file, data, names):
    """pickle.dump({names[i]: data[i]}) to path/file (BasicFunctionsSJR.py:25-50)."""
    mkdir(path)
    with open(os.path.join(path, file), 'wb') as f:
        pickle.dump({names[i]: data[i] for i in range(len(names))}, f)


def load_pr(path_file, names=None):
    """whole dict, one entry (names is a str) or a tuple of entries; False when the file does not exist
    (BasicFunctionsSJR.py:53-89)."""
    if not os.path.isfile(path_file):
        return False
    with open(path_file, 'rb') as f:
        data = pickle.load(f)
    if names is None:
        return data
    if isinstance(names, str):
        return data[names]
    return tuple(data[n] for n in names)


def arg_find_array(arg, n=1, which='first'):
    """index of the n-th True entry counted from the front ('first') or the back ('last')"""
    import numpy as np
    idx = np.nonzero(np.asarray(arg).reshape(-1))[0]
    if idx.size < n:
        return -1
    return int(idx[n - 1]) if which == 'first' else int(idx[-n])


def empty_list(n, content=None):
    import copy
    return [copy.copy(content) for _ in range(n)]
