"""Pretty-print the key numbers of a bench.py JSON line.  usage: python scripts/show_bench.py FILE"""
import json
import sys

for l in open(sys.argv[1]):
    l = l.strip()
    if not l.startswith("{"):
        continue
    d = json.loads(l)

    def show(d, ind=0, maxlen=110):
        for k, v in d.items():
            if isinstance(v, dict):
                print(" " * ind + k + ":")
                show(v, ind + 2)
            else:
                print(" " * ind + f"{k}: {str(v)[:maxlen]}")

    show(d)
