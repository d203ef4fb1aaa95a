#!/usr/bin/env python3
"""Build kernel-variant copies of librabe_b200.so under build/variants/ (development aid for
occupancy sweeps on the GPU box: RABE_B200_LIB=build/variants/<name>.so python bench.py ...)."""
import os, sys, subprocess
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rabe_b200 import build as rb

VARIANTS = {
    "noco": ["RB_COOP_PAIRING=0"],
    "co2": ["RB_CO_MINB=2"],
    "co3": ["RB_CO_MINB=3"],
    "co4": ["RB_CO_MINB=4"],
    "co5": ["RB_CO_MINB=5"],
    "co6": ["RB_CO_MINB=6"],
    "co8b64": ["RB_CO_MINB=8", "RB_CO_BLOCK=64"],
    "fe64": ["RB_CO_FE_BLOCK=64"],
    "fe32": ["RB_CO_FE_BLOCK=32"],
    "item": ["RB_DEC_ITEM=1"],
    "item64": ["RB_DEC_ITEM=1", "RB_CO_BLOCK=64", "RB_CO_FE_BLOCK=64"],
    "ml256fe128": ["RB_CO_BLOCK=256", "RB_CO_FE_BLOCK=128"],
    "ml128fe256": ["RB_CO_BLOCK=128", "RB_CO_FE_BLOCK=256"],
    "ml256fe256": ["RB_CO_BLOCK=256", "RB_CO_FE_BLOCK=256"],
    "cob64": ["RB_CO_BLOCK=64"],
    "cob256": ["RB_CO_BLOCK=256"],
    "cob32": ["RB_CO_BLOCK=32"],
    "g1b3": ["RB_G1_MINB=3"],
    "g1b4": ["RB_G1_MINB=4"],
    "g1b5": ["RB_G1_MINB=5"],
    "g1b6": ["RB_G1_MINB=6"],
    "cmp": ["RB_CO_MULFP_NOINLINE=1", "RB_STEP_NOINLINE=1", "RB_COMPACT=1"],
    "cmp3": ["RB_CO_MULFP_NOINLINE=1", "RB_STEP_NOINLINE=1", "RB_COMPACT=1", "RB_CO_MINB=3"],
    "mfn": ["RB_CO_MULFP_NOINLINE=1"],
    "mfn_sn": ["RB_CO_MULFP_NOINLINE=1", "RB_STEP_NOINLINE=1"],
    "g1m24": ["RB_G1_M=24"],
    "g1m32": ["RB_G1_M=32"],
    "sn": ["RB_STEP_NOINLINE=1"],
    "sn3": ["RB_STEP_NOINLINE=1", "RB_CO_MINB=3"],
    "sn4": ["RB_STEP_NOINLINE=1", "RB_CO_MINB=4"],
    "m3": ["RB_PAIR_MINB=3"],
    "m4": ["RB_PAIR_MINB=4"],
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
