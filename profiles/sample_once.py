"""A few reverse steps of the sampler on N images (driver for ncu). usage: python profiles/sample_once.py <N> <t_start>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import igdm_b200  # noqa
from igdm_b200 import api, capi, tables
N, t0 = int(sys.argv[1]), int(sys.argv[2])
h = capi.Handle(T=500, precision=1)
beta, _, acum = tables.beta_schedule(500)
h.set_tables(beta, acum, tables.embedding_table(500))
h.set_weights(api.SimpleUNet.load().arrays)
h.set_option("sample_chunk", N)
h.set_option("use_graph", 0)
h.sample_device(N, seed=1, first_index=0, t_start=t0)
print("ok", h.counter("launches"))
