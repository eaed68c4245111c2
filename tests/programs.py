"""Programs used across the parity tests.  They follow the reference's own end-to-end tests
(/root/reference/src/tests/mod.rs:12-309, tests/comparisons.rs) and example programs (src/examples/*.rs)."""
from distaff_b200 import hostvm


def merkle_example(depth, po):
    """examples/merkle.rs:60-108 (authentication path drawn from field::prng_vector with the example's two seeds)"""
    s1 = bytes([1, 2, 3] + [0] * 29)
    s2 = bytes([4, 5, 6] + [0] * 29)
    p0, p1 = po.prng_vector(s1, depth), po.prng_vector(s2, depth)
    leaf_index = p0[0] % (2 ** (depth - 1))
    a, b = [p0[0]], [p1[0]]
    index = leaf_index + 2 ** (depth - 1)
    for i in range(1, depth):
        a += [0, p0[i]]
        b += [index & 1, p1[i]]
        index >>= 1
    for i in range(1, depth):
        a.append(p0[i])
        b.append(p1[i])
    return hostvm.execute(hostvm.merkle_program(depth, leaf_index), secret_a=a, secret_b=b, num_outputs=4)


def merkle_paths(depth, count, po):
    """`count` Merkle authentication paths of length `depth` verified back to back with the program of examples/merkle.rs:41-57
    (smpath then pmpath, both Rescue-based); seeds follow the example's, with the path number in byte 3.  The example itself cannot
    reach BASELINE's 2^14 steps: its index arithmetic (`usize::pow(2, n - 1)`, examples/merkle.rs:75,106) caps the depth at 64
    (2^12 steps), and pmpath's binary decomposition needs 2^(depth-1) < M.  Four depth-64 paths give the 2^14-step, Rescue-dominated
    trace that config names."""
    a_all, b_all, blocks = [], [], []
    for c in range(count):
        s1 = bytes([1, 2, 3, c] + [0] * 28)
        s2 = bytes([4, 5, 6, c] + [0] * 28)
        p0, p1 = po.prng_vector(s1, depth), po.prng_vector(s2, depth)
        leaf_index = p0[0] % (2 ** (depth - 1))
        a, b = [p0[0]], [p1[0]]
        index = leaf_index + 2 ** (depth - 1)
        for i in range(1, depth):
            a += [0, p0[i]]
            b += [index & 1, p1[i]]
            index >>= 1
        for i in range(1, depth):
            a.append(p0[i])
            b.append(p1[i])
        a_all += a
        b_all += b
        blocks.append(f"read.ab dup.2 smpath.{depth} swap.2 push.{leaf_index} roll.4 swap swap.2 pmpath.{depth}")
    return hostvm.execute("begin " + " drop.4 ".join(blocks) + " end", secret_a=a_all, secret_b=b_all, num_outputs=4)


def small_programs():
    """name -> ExecutionTrace ; all have 2^6..2^9 steps so the CPU oracle proves each in well under a second"""
    P = {}
    P["fib13"] = hostvm.fibonacci(13)                                              # BASELINE configs[0]
    P["fib_span"] = hostvm.execute("begin swap dup.2 drop add swap dup.2 drop add swap dup.2 drop add end",
                                   public_inputs=[1, 0])                           # tests/mod.rs:12-29
    P["stack_ops"] = hostvm.execute("begin swap swap.2 swap.4 roll.4 roll.8 pad.2 drop.2 dup dup.2 dup.4 drop.8 end",
                                    public_inputs=[1, 2, 3, 4, 5, 6, 7, 8], num_outputs=8)
    P["logic"] = hostvm.execute("begin not and or end", public_inputs=[1, 1, 0, 1], num_outputs=2)   # tests/mod.rs logic_operations
    P["arith"] = hostvm.execute("begin add mul inv neg push.3 sub push.9 div end", public_inputs=[2, 3, 4], num_outputs=1)
    P["eq"] = hostvm.execute("begin eq swap.2 ne and end", public_inputs=[5, 5, 9, 6, 7], num_outputs=1)
    P["cmp"] = hostvm.execute("begin push.5 push.11 gt.8 push.7 push.3 lt.8 and end", num_outputs=1)      # tests/comparisons.rs
    P["rc"] = hostvm.execute("begin push.200 rc.8 push.300 rc.8 end", num_outputs=2)
    P["choose"] = hostvm.execute("begin choose swap.2 choose.2 end", public_inputs=[3, 4, 1, 5, 6, 7, 0, 8], num_outputs=3)
    P["hash"] = hostvm.execute("begin pad.2 hash.2 end", public_inputs=[5, 6], num_outputs=2)             # tests/mod.rs hash_operations
    P["if_else"] = hostvm.execute("begin push.3 push.5 read if.true add else mul end end", secret_a=[1], num_outputs=1)
    P["if_else0"] = hostvm.execute("begin push.3 push.5 read if.true add else mul end end", secret_a=[0], num_outputs=1)
    P["nested"] = hostvm.execute("begin push.2 block push.3 block push.4 add end mul end add end", public_inputs=[1], num_outputs=1)
    P["collatz3"] = hostvm.collatz(3)                                                                  # loop + switch + isodd
    # wider register files: user stack deeper than 16, context depth 3, nested loops (loop depth 2)
    P["deep_stack"] = hostvm.execute("begin " + " ".join(f"push.{i + 1}" for i in range(22)) + " add mul swap.4 roll.8 dup.4 drop.8 end",
                                     public_inputs=[9, 8, 7], num_outputs=4)
    P["deep_ctx"] = hostvm.execute("begin push.1 block push.2 block push.3 block push.4 add end add end add end end", num_outputs=1)
    P["nested_loops"] = hostvm.execute(
        "begin push.2 push.1 while.true push.3 push.1 while.true push.1 neg add dup push.0 ne end drop push.1 neg add dup push.0 ne end end",
        num_outputs=1)
    return P


def wide_program():
    """65 registers (ctx 15, loop 3, stack 32): every extended-trace row is 1040 bytes = two BLAKE3 chunks + a parent node
    (trace_table.rs:174-185 with w > 64).  Kept out of small_programs(): the golden file and the GPU parametrisations are per name."""
    from distaff_b200 import hostvm
    loop = ("push.2 push.1 while.true push.2 push.1 while.true push.2 push.1 while.true push.1 neg add dup push.0 ne end drop "
            "push.1 neg add dup push.0 ne end drop push.1 neg add dup push.0 ne end drop")
    src = "begin " + " ".join(f"push.{i + 1}" for i in range(26)) + " " + "block push.1 add " * 15 + "end " * 15 + " " + loop + " end"
    return hostvm.execute(src, num_outputs=2)
