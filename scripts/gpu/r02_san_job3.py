import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import plaac_b200
from tests import synth
n = int(sys.argv[1])
lc, lo = synth.long_proteins(seed=1009, lengths=(n,))
sc = plaac_b200.Scorer(); sc.set_long_path(1024)
s = sc.score(lc, lo)
print("summary ok", n, sc.stats().long_proteins)
