"""Pins oracle/air.h against the reference's TraceState tests (/root/reference/src/stark/trace/trace_state.rs:500-579)
and checks that every transition constraint vanishes on genuine VM traces (the property the reference asserts at
/root/reference/src/stark/constraints/evaluator.rs:152-157)."""
import numpy as np

M = 2**128 - 45 * 2**40 + 1


def _flags(po, row, cd=1, ld=0, sd=2):
    a = po.fvec(row)
    out = np.zeros((46, 2), dtype=np.uint64)
    po.lib().or_op_flags(a.ctypes.data, cd, ld, sd, out.ctypes.data)
    v = po.ints(out)
    return v[:8], v[8:40], v[40:44], v[44], v[45]


def test_op_flags_reference_vectors(po):
    cf, ld, hd, begin, noop = _flags(po, [101, 1, 2, 3, 4, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 15, 16, 17])
    assert cf == [1, 0, 0, 0, 0, 0, 0, 0] and ld == [0] * 32 and hd == [0, 0, 0, 0] and (begin, noop) == (1, 0)
    cf, ld, hd, begin, noop = _flags(po, [101, 1, 2, 3, 4, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 15, 16, 17])
    assert cf == [0] * 7 + [1] and ld == [0] * 31 + [1] and hd == [0, 0, 0, 1] and (begin, noop) == (0, 1)
    cf, ld, hd, begin, noop = _flags(po, [101, 1, 2, 3, 4, 1, 0, 0, 1, 0, 0, 0, 0, 1, 0, 15, 16, 17])
    assert cf == [0, 1, 0, 0, 0, 0, 0, 0] and ld == [0, 1] + [0] * 30 and hd == [0, 1, 0, 0] and (begin, noop) == (0, 0)
    cf, ld, hd, _, _ = _flags(po, [101, 1, 2, 3, 4, 1, 1, 0, 1, 1, 0, 0, 0, 0, 1, 15, 16, 17])
    assert cf == [0, 0, 0, 1, 0, 0, 0, 0] and ld == [0, 0, 0, 1] + [0] * 28 and hd == [0, 0, 1, 0]


def test_ld_flag_quirk_uses_cf_bit(po):
    # trace_state.rs:301: ld_op_flags[2] = (1 - ld0) * cf_bits[1]
    row = [0, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 1, 1, 0, 0, 0]   # cf=[0,1,0], ld=[0,1,0,0,0], hd=[1,1]
    _, ld, _, _, _ = _flags(po, row)
    assert ld[2] == 1
    row[6] = 0                                                     # cf=[0,0,0]: the flag for ld index 2 disappears
    _, ld, _, _, _ = _flags(po, row)
    assert ld[2] == 0


def _rows(trace, step):
    return trace.registers[:, step, :].copy()


def test_transition_constraints_vanish_on_vm_traces(po):
    from distaff_b200 import hostvm
    programs = [
        hostvm.fibonacci(6),
        hostvm.execute("begin push.3 push.5 add push.7 mul dup inv mul not not end", num_outputs=1),
        hostvm.collatz(3),
        hostvm.execute("begin pad.2 hash.2 end", public_inputs=[5, 6], num_outputs=2),
        hostvm.execute("begin push.5 push.11 gt.8 push.7 push.3 lt.8 and end", num_outputs=1),
        hostvm.execute("begin read.ab read.ab swap.2 drop choose push.9 eq end", secret_a=[3, 1], secret_b=[4, 7], num_outputs=1),
    ]
    for tr in programs:
        n = tr.length
        out = np.zeros((64, 2), dtype=np.uint64)
        for step in range(n - 1):
            cur, nxt = _rows(tr, step), _rows(tr, step + 1)
            k = po.lib().or_eval_transition_raw(cur.ctypes.data, nxt.ctypes.data, tr.ctx_depth, tr.loop_depth, tr.stack_depth,
                                                n, step * 8, out.ctypes.data)
            vals = po.ints(out[:k])
            assert all(v == 0 for v in vals), (step, [i for i, v in enumerate(vals) if v])
        # a corrupted next row must violate something
        cur, nxt = _rows(tr, 3), _rows(tr, 4)
        nxt[tr.width - 1, 0] ^= np.uint64(1)
        k = po.lib().or_eval_transition_raw(cur.ctypes.data, nxt.ctypes.data, tr.ctx_depth, tr.loop_depth, tr.stack_depth,
                                            n, 3 * 8, out.ctypes.data)
        assert any(v != 0 for v in po.ints(out[:k]))
