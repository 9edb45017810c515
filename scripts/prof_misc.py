"""Profiling targets: 'var' = fold_vruns on the scaled configs[3] log, 'inc' = K6 micro-batches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
what = sys.argv[1]
if what == "var":
    sys.argv = ["config4.py", "312500", "10000000", "no"]
    exec(open(os.path.join(os.path.dirname(__file__), "config4.py")).read())
else:
    sys.argv = ["config5.py", "20"]
    exec(open(os.path.join(os.path.dirname(__file__), "config5.py")).read())
