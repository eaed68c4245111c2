#!/usr/bin/env python3
"""Extracts the literal test vectors of the reference's decoder unit tests into a JSON fixture.

Source: /root/reference/src/stark/constraints/decoder/flow_ops.rs (mod tests, lines 170-455): every case builds two TraceStates with
`new_state(step, flow_op, sponge, ctx_stack, loop_stack)`, calls one `enforce_<op>(&mut evaluations, &state1, &state2, 1)` and compares
with a literal vector (`are_equal(a, b)` = a - b in the field, constraints/utils.rs:24-26).
Run in the build container (the reference is not present on the GPU box):
    python tests/golden/make_ref_decoder_cases.py  ->  tests/golden/ref_decoder_flow_cases.json
The consumer (tests/test_oracle_air.py) feeds the same two states, with the op bits of <op>, to the oracle's decoder evaluation."""
import json
import os
import re

M = 2**128 - 45 * 2**40 + 1
SRC = "/root/reference/src/stark/constraints/decoder/flow_ops.rs"
HERE = os.path.dirname(os.path.abspath(__file__))


def ints(s):
    s = s.strip()
    return [int(x) for x in s.split(",")] if s else []


def expr(tok):
    tok = tok.strip()
    m = re.fullmatch(r"are_equal\((\d+),\s*(\d+)\)", tok)
    if m:
        return (int(m.group(1)) - int(m.group(2))) % M
    return int(tok)


def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        if ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def main():
    text = open(SRC).read()
    text = text[text.index("mod tests"):]
    state_re = re.compile(r"let state(\d) = new_state\((\d+), FlowOps::(\w+),\s*&\[([^\]]*)\], &\[([^\]]*)\], &\[([^\]]*)\]\);")
    call_re = re.compile(r"super::enforce_(\w+)\(&mut evaluations, &state1, &state2, (\d+)\);")
    want_re = re.compile(r"assert_eq!\(vec!\[(.*)\], evaluations\);")
    cases, st = [], {}
    fn = None
    for line in text.splitlines():
        m = state_re.search(line)
        if m:
            st[m.group(1)] = {"step": int(m.group(2)), "flow_op": m.group(3), "sponge": ints(m.group(4)), "ctx": ints(m.group(5)), "loop": ints(m.group(6))}
            continue
        m = call_re.search(line)
        if m:
            fn = (m.group(1), int(m.group(2)))
            continue
        m = want_re.search(line)
        if m and fn:
            cases.append({"op": fn[0], "op_flag": fn[1], "state1": st["1"], "state2": st["2"], "expected": [str(expr(t)) for t in split_top(m.group(1))]})
            fn = None
    assert len(cases) >= 25, len(cases)
    with open(os.path.join(HERE, "ref_decoder_flow_cases.json"), "w") as f:
        f.write('{"source": "src/stark/constraints/decoder/flow_ops.rs mod tests", "cases": [\n' + ",\n".join(json.dumps(c) for c in cases) + "\n]}\n")
    print(len(cases), "cases;", sorted({c["op"] for c in cases}))


if __name__ == "__main__":
    main()
