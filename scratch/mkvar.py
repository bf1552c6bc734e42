"""Builds tuning variants (retrofire_b200/_variants/<name>.so) in parallel: python scratch/mkvar.py name:K=V,K=V ..."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from retrofire_b200 import build as b
import concurrent.futures as cf
V = {}
for a in sys.argv[1:]:
    n, _, d = a.partition(':')
    V[n] = dict(kv.split('=') for kv in d.split(',') if kv)
with cf.ThreadPoolExecutor(6) as ex:
    list(ex.map(lambda kv: b.build_variant(kv[0], kv[1]), V.items()))
print("built", list(V))
