"""Experiment builds of the CUDA library: `python tools/build_variants.py name:MACRO=V,MACRO=V ...` -> build/variants/<name>.so
(loaded with SDFIBM_B200_LIB=build/variants/<name>.so)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdfibm_b200 import build as b
out_dir = os.path.join(b.ROOT, "build", "variants")
os.makedirs(out_dir, exist_ok=True)
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    out = os.path.join(out_dir, name + ".so")
    b.build(force=True, defines=[d for d in defs.split(",") if d], out=out)
    print(out)
