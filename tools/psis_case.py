import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
import viabel_b200 as vb
from _problems import psis_case
name = sys.argv[1]
lw = psis_case(name)
out, k, ti, tr = vb.psislw(lw, return_tail=True)
print(name, k, out[:3], len(ti))
