"""Build experimental variants of the library (compile-time switches) into build/variants/ for A/B timing on the
GPU box: python tools/variants.py NAME=DEF1,DEF2 ...   then   P2B_LIB=build/variants/NAME.so python tools/perm_bench.py"""
import importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "plonky2-gpu_b200", "build.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
for arg in sys.argv[1:]:
    name, _, defs = arg.partition("=")
    out = os.path.join(ROOT, "variants", name + ".so")
    b.build(force=True, defines=[d for d in defs.split(",") if d], out=out)
    print("built", out)
