import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import distaff_b200 as dg
from oracle import pyoracle as po
from tests import programs
P = programs.small_programs()
tr = P["collatz3"]
for rep in range(2):
    for opts in [(32, 50, 20), (16, 30, 8), (64, 20, 12), (128, 10, 0), (256, 5, 4), (256, 5, 4), (128, 10, 0), (256, 50, 4)]:
        ref = po.prove(tr.registers, tr.ctx_depth, tr.loop_depth, tr.public_inputs, tr.outputs, ext=opts[0], num_queries=opts[1], grinding=opts[2])
        p = dg.prove(tr, dg.ProofOptions(*opts))
        same = p.bytes == ref.proof
        msg = ""
        if not same:
            diff = [i for i in range(min(len(p.bytes), len(ref.proof))) if p.bytes[i] != ref.proof[i]]
            msg = f"len {len(p.bytes)} vs {len(ref.proof)} first diffs {diff[:6]} n={len(diff)} roots {p.trace_root == ref.digest('trace_root')} {p.constraint_root == ref.digest('constraint_root')} nonce {p.pow_nonce} {ref.u64s('pow_nonce')[0]}"
        print(opts, same, msg, flush=True)
