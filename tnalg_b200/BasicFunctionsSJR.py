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


# ---- small utilities used by the reference's scripts and parameter printing (BasicFunctionsSJR.py:16-22,107-163,
# 204-240,369-470); plain text instead of termcolor ----
def info_contact():
    return {'name': 'S.J. Ran', 'email': 'ranshiju10@mail.s ucas.ac.cn', 'affiliation': 'ICFO – The Institute of Photonic Sciences'}


def search_file(path, exp):
    import re
    pattern = re.compile(exp)
    return [os.path.join(path, x) for x in os.listdir(path) if re.match(pattern, x)]


def output_txt(x, filename='data'):
    import numpy as np
    np.savetxt(filename + '.txt', x)


def sort_list(a, order):
    """[a[i] for i in order]"""
    return [a[i] for i in order]


def remove_element_from_list(x, element):
    return [a for a in x if a != element]


def arg_find_list(x, target, n=1, which='first'):
    """positions of the first (or last) n occurrences of target in a list; positions of the 'last' search refer to the
    original order"""
    x = list(x)
    idx = [i for i, a in enumerate(x) if a == target]
    if which == 'last':
        idx = idx[::-1]
    return idx[:n]


def print_dict(a, keys=None, welcome='', style_sep=': ', color='white', end='\n'):
    express = welcome
    if keys is None:
        for n in a:
            express += str(n) + style_sep + str(a[n]) + end
    elif isinstance(keys, str):
        express += keys.capitalize() + style_sep + str(a[keys])
    else:
        keys = list(keys)
        for i, n in enumerate(keys):
            express += n.capitalize() + style_sep + str(a[n])
            if i != len(keys) - 1:
                express += end
    print(express)
    return express


def print_error(string, if_trace_stack=True):
    print(string)
    if if_trace_stack:
        import traceback
        traceback.print_stack(limit=3)


def print_sep(info='', style='=', length=40, color='cyan'):
    if info == '':
        mes = style * (length * 2)
    else:
        l_new = length * 2 - 2 - len(info)
        dl = l_new % 2
        l_new = max(int(l_new / 2), 0)
        mes = style * l_new + ' ' + info + ' ' + style * ((l_new + dl) * (l_new > 0))
    print(mes)
    return mes


def print_options(options, start=None, welcome='', style_sep=': ', end='    ', color='cyan', quote=None):
    if start is None:
        start = list(range(len(options)))
    parts = []
    for i, o in enumerate(options):
        parts.append(str(start[i]) + style_sep + (o if quote is None else quote + o + quote))
    message = welcome + end.join(parts)
    print(message)
    return message
