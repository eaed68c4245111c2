"""Python face of the host-side VM stand-in (`distaff_b200/hostvm/vm.cpp`).

Mirrors the host half of `distaff::execute` (/root/reference/src/lib.rs:30-59): compile assembly, run the VM, return the
column-major register traces plus what `stark::prove` needs (ctx/loop depth, public inputs, outputs, program hash).
This is an input generator for tests / smoke / bench; the prove hot path never calls it.
"""
import ctypes
import os
import numpy as np

from .. import felt

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libdistaff_vm.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'` (or make -C distaff_b200/hostvm)")
        lib = ctypes.CDLL(path)
        lib.vm_execute.restype = ctypes.c_void_p
        lib.vm_execute.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32]
        lib.vm_error.restype = ctypes.c_char_p
        lib.vm_error.argtypes = [ctypes.c_void_p]
        for name in ("vm_width", "vm_ctx_depth", "vm_loop_depth", "vm_stack_depth"):
            getattr(lib, name).restype = ctypes.c_uint32
            getattr(lib, name).argtypes = [ctypes.c_void_p]
        lib.vm_length.restype = ctypes.c_uint64
        lib.vm_length.argtypes = [ctypes.c_void_p]
        lib.vm_program_hash.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.vm_copy_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.vm_free.argtypes = [ctypes.c_void_p]
        _LIB = lib
    return _LIB


class ExecutionTrace:
    """What crosses the `stark::prove` seam (/root/reference/src/lib.rs:62)."""

    def __init__(self, registers, ctx_depth, loop_depth, stack_depth, program_hash, public_inputs, outputs):
        self.registers = registers          # (w, n, 2) uint64, column-major register traces
        self.ctx_depth = ctx_depth
        self.loop_depth = loop_depth
        self.stack_depth = stack_depth
        self.program_hash = program_hash    # 32 bytes
        self.public_inputs = public_inputs  # list[int]
        self.outputs = outputs              # list[int]

    @property
    def width(self):
        return self.registers.shape[0]

    @property
    def length(self):
        return self.registers.shape[1]


def execute(source, public_inputs=(), secret_a=(), secret_b=(), num_outputs=1):
    lib = _lib()
    pub = felt.from_ints(public_inputs)
    sa = felt.from_ints(secret_a)
    sb = felt.from_ints(secret_b)
    h = lib.vm_execute(source.encode(), pub.ctypes.data, len(pub), sa.ctypes.data, len(sa), sb.ctypes.data, len(sb))
    try:
        err = lib.vm_error(h)
        if err:
            raise RuntimeError("vm: " + err.decode())
        w, n = lib.vm_width(h), lib.vm_length(h)
        regs = np.empty((w, n, 2), dtype=np.uint64)
        lib.vm_copy_trace(h, regs.ctypes.data)
        ph = ctypes.create_string_buffer(32)
        lib.vm_program_hash(h, ph)
        cd, ld, sd = lib.vm_ctx_depth(h), lib.vm_loop_depth(h), lib.vm_stack_depth(h)
    finally:
        lib.vm_free(h)
    assert num_outputs <= 8
    stack_start = 15 + cd + ld
    # outputs = top of the user stack at the last step, padded with zeros to the state's min depth 8 (lib.rs:45-46)
    last = [0] * 8
    for i in range(min(sd, 8)):
        last[i] = felt.to_ints(regs[stack_start + i, n - 1])[0]
    outputs = last[:num_outputs]
    return ExecutionTrace(regs, cd, ld, sd, ph.raw, list(public_inputs), outputs)


# ---- example programs (inputs of BASELINE.json's configs; /root/reference/src/examples/*.rs) --------------------------
def fibonacci_program(n):
    """examples/fibonacci.rs:33-42 ; inputs [1, 0], 1 output"""
    return f"begin repeat.{n - 1} swap dup.2 drop add end end"


def fibonacci(n):
    return execute(fibonacci_program(n), public_inputs=[1, 0], num_outputs=1)


COLLATZ_SOURCE = """
begin
    pad read dup push.1 ne
    while.true
        swap push.1 add swap dup isodd.128
        if.true
            push.3 mul push.1 add
        else
            push.2 div
        end
        dup push.1 ne
    end
    swap
end"""


def collatz(start):
    """examples/collatz.rs:10-23 ; secret tape A = [start], 1 output (number of steps)"""
    return execute(COLLATZ_SOURCE, secret_a=[start], num_outputs=1)


def merkle_program(depth, index):
    """examples/merkle.rs:41-57"""
    return f"begin read.ab dup.2 smpath.{depth} swap.2 push.{index} roll.4 swap swap.2 pmpath.{depth} end"
