"""Tiny key/value logger with the call surface the learner uses from the reference's baselines/logger.py
(logkv :222, logkv_mean :231, dumpkvs :246, info :277, get_dir :300, configure :372): human-readable table on
stdout and, when a directory is configured (or $OPENAI_LOGDIR is set), `progress.csv` (logger.py:102-136).
"""
import csv
import os
import sys
from collections import OrderedDict, defaultdict

_kvs = OrderedDict()
_counts = defaultdict(int)
_dir = None
_csv_keys = None
_quiet = False


def configure(dir=None, format_strs=None, quiet=False):
    global _dir, _csv_keys, _quiet
    _dir = dir or os.environ.get("OPENAI_LOGDIR")
    _csv_keys = None
    _quiet = quiet
    if _dir:
        os.makedirs(_dir, exist_ok=True)


def get_dir():
    return _dir


def logkv(key, val):
    _kvs[key] = val


def logkv_mean(key, val):
    old, cnt = _kvs.get(key, 0.0), _counts[key]
    _kvs[key] = old * cnt / (cnt + 1) + val / (cnt + 1)
    _counts[key] = cnt + 1


def logkvs(d):
    for k, v in d.items():
        logkv(k, v)


def getkvs():
    return _kvs


def info(*args):
    if not _quiet:
        print(*args, file=sys.stdout, flush=True)


log = info


def dumpkvs():
    global _csv_keys
    d = OrderedDict(_kvs)
    if not _quiet and d:
        kw = max(len(str(k)) for k in d)
        vals = {k: (f"{v:.6g}" if isinstance(v, float) else str(v)) for k, v in d.items()}
        vw = max(len(v) for v in vals.values())
        line = "-" * (kw + vw + 7)
        print(line)
        for k in sorted(d):
            print(f"| {k:<{kw}} | {vals[k]:<{vw}} |")
        print(line, flush=True)
    if _dir and d:
        path = os.path.join(_dir, "progress.csv")
        if _csv_keys is None:
            _csv_keys = sorted(d)
            with open(path, "w", newline="") as f:
                csv.writer(f).writerow(_csv_keys)
        with open(path, "a", newline="") as f:
            csv.writer(f).writerow([d.get(k, "") for k in _csv_keys])
    _kvs.clear()
    _counts.clear()
    return d
