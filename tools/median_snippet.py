"""Two frames through a 7x7 windowed context at 1080p: the command profiles/r02_median7_ncu.txt was captured on
(ncu --set full -k regex:spatial_median -c 1 python tools/median_snippet.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, dips_b200
with dips_b200.Context(1920,1080,1,0,32,spatial_window=7) as c:
    f=np.random.default_rng(1).integers(0,256,(1920*1080*4,),dtype=np.uint8)
    c.push_frame(f); c.push_frame(f)
print("ok")
