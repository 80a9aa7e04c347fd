"""Per-role cycle breakdown of the tcgen05 conv kernel (clock64 instrumentation inside the kernel).
usage: python profiles/role_profile.py [n_images]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import igdm_b200  # noqa: E402,F401
from igdm_b200 import api, capi, tables  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
h = capi.Handle(T=500, precision=capi.PREC_FP16)
beta, _, acum = tables.beta_schedule(500)
h.set_tables(beta, acum, tables.embedding_table(500))
h.set_weights(api.SimpleUNet.load().arrays)
if len(sys.argv) > 2:
    h.set_option("tc_tma_store", int(sys.argv[2]))
if len(sys.argv) > 3:
    h.set_option("tc_pair", int(sys.argv[3]))
for name in ("conv_l2", "conv_l9", "conv_l4", "conv_l3", "up2"):
    ms, by, fl = h.time_kernel(name, n, 10)
    h.set_option("tc_role_profile", 1)
    h.time_kernel(name, n, 1)   # 3 warm-up + 1 timed launches accumulate into the counters
    d = h.debug_fetch("tc_roles").reshape(512, 8)
    h.set_option("tc_role_profile", 0)
    live = d[d[:, 6] > 0]
    tiles = live[:, 6].sum()
    per = live.sum(axis=0) / tiles                 # slots 4..6 come from epilogue set 0 (every other tile of the CTA)
    iss = live[live[:, 3] > 0]                     # CTAs whose issuer warp 1 worked (pair mode: leaders only)
    per[1:4] = iss[:, 1:4].sum(axis=0) / iss[:, 0].sum()      # per tile issued by that warp (slot 0 = its tile count)
    per[7] = live[:, 7].sum() / (2 * tiles)        # kernel cycles per tile of the CTA (both epilogue sets)
    print(f"{name}: {ms*1e3:.1f} us {(fl or 0)/ms/1e9:.0f} TFLOP/s | ctas={len(live)} tiles/launch={tiles/4:.0f} | cycles per tile: "
          f"mma_wait_acc_empty={per[1]:.0f} mma_wait_a_full={per[2]:.0f} mma_issue={per[3]:.0f} "
          f"epi_wait_acc_full={per[4]:.0f} epi_busy={per[5]:.0f} kernel_cycles_per_tile={per[7]:.0f}")
