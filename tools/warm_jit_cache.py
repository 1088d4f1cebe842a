#!/usr/bin/env python
"""Fill C4B_JIT_CACHE_DIR with the specialised kernels of the shipped models -- the device
counterpart of running the reference's bootstrapper once (src/c4/bootstrapper.c).  Needs no GPU
(NVRTC cross-compiles for sm_100a).  BSDP's derived models are compiled on first use by the
run itself and land in the same directory.
usage: python tools/warm_jit_cache.py <cache-dir> [model ...]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) < 2:
    sys.exit(__doc__)
os.makedirs(sys.argv[1], exist_ok=True)
os.environ["C4B_JIT_CACHE_DIR"] = os.path.abspath(sys.argv[1])
from exonerate_b200 import load_library          # noqa: E402
from exonerate_b200.models import host_model     # noqa: E402

lib = load_library()
names = sys.argv[2:] or ["ungapped", "affine:local", "affine:global", "est2genome", "protein2genome",
                         "coding2coding"]
for name in names:
    model, _ = host_model(name, query_is_protein=name.startswith("protein2"))
    t0 = time.time()
    for threads in (128, 256, 512):
        for mode in (0, 1, 2):
            if lib.c4b_model_specialise(C.byref(model), mode, threads, None):
                sys.exit("%s: %s" % (name, lib.c4b_last_error().decode()[:2000]))
    print("%-16s 3 CTA sizes x 3 modes in %.1f s" % (name, time.time() - t0))
print("%d cubins in %s" % (len(os.listdir(sys.argv[1])), sys.argv[1]))
