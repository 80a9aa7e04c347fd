import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import igdm_b200  # noqa
from igdm_b200 import api, capi, tables
h = capi.Handle(T=500, precision=1)
beta, _, acum = tables.beta_schedule(500)
h.set_tables(beta, acum, tables.embedding_table(500))
h.set_weights(api.SimpleUNet.load().arrays)
n, graph, tstart, fuse = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
h.set_option("use_graph", graph); h.set_option("fuse_final", fuse)
h.sample_device(n, seed=1, t_start=tstart); print("sample", n, graph, tstart, fuse, "ok", flush=True)
