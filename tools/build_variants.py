#!/usr/bin/env python3
"""Build kernel-variant copies of librabe_b200.so under build/variants/ (development aid for
occupancy sweeps on the GPU box: RABE_B200_LIB=build/variants/<name>.so python bench.py ...)."""
import os, sys, subprocess
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rabe_b200 import build as rb

VARIANTS = {
    "m3": ["RB_PAIR_MINB=3"],
    "m4": ["RB_PAIR_MINB=4"],
    "m8b64": ["RB_PAIR_MINB=8", "RB_ML_BLOCK=64", "RB_FE_BLOCK=64"],
    "m5b64": ["RB_PAIR_MINB=5", "RB_ML_BLOCK=64", "RB_FE_BLOCK=64"],
}

def main():
    names = sys.argv[1:] or list(VARIANTS)
    rb.build()                      # host objects + the product library
    os.makedirs(os.path.join(ROOT, "build", "variants"), exist_ok=True)
    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(lambda n: rb.build(defines=VARIANTS[n], out=os.path.join(ROOT, "build", "variants", n + ".so")), names))
    print("built", names)

if __name__ == "__main__":
    main()
