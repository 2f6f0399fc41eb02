import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, strkit_b200 as sb
from tests.helpers import families_to_batch
eng = sb.Engine(); p = sb.RepeatCountParams("repalign", 50, 3, 1)
b = families_to_batch([("CAGCAG", "CAGCAG", "AC", "GT")], est=[(1 << 22) - 1])
print(b.est_cn, b.motif_len, b.lens)
try:
    print(eng.count_reads(b, p))
except Exception as e:
    print("raised", e)
