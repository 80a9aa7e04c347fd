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
h.set_option("use_graph", 0)
n = int(sys.argv[1]); 
for name in sys.argv[2:]:
    if name.startswith("conv"):
        print(name, n, h.time_kernel(name, n, 2), flush=True)
    else:
        h.sample_device(n, seed=1, t_start=3); print("sample", n, "ok", flush=True)
