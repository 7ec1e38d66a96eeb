"""The run-time specialisation path builds without a GPU: generate + NVRTC (sm_100a) for the
passes of real compiled programs.  No kernel is launched here."""
import ctypes as C

import pytest

from quantum_simulations_b200 import _lib as L, workloads as W
from quantum_simulations_b200.kernel.cuda_dense import compile_circuit


def _build(step, dtype=L.QSV_C128):
    n, log = C.c_size_t(), C.create_string_buffer(1 << 14)
    rc = L.load().qsv_jit_build_pass(C.byref(step.desc), step.ops, dtype, C.byref(n), log, len(log))
    return rc, n.value, log.value.decode(errors="replace")


@pytest.mark.parametrize("circuit", ["random_1q_cz", "qft", "random_mixed"])
def test_every_ring_pass_specialises(circuit):
    cd = {"random_1q_cz": lambda: W.random_1q_cz(16, 20, 1234), "qft": lambda: W.qft(14),
          "random_mixed": lambda: W.random_mixed(14, 200, 3)}[circuit]()
    prog = compile_circuit(cd)
    assert prog.passes
    for step in prog.passes[:3]:
        rc, size, log = _build(step)
        if rc == L.QSV_EIO:
            pytest.skip(f"NVRTC not installed: {log}")
        assert rc == 0 and size > 10000, log


def test_complex64_passes_specialise_too():
    prog = compile_circuit(W.random_1q_cz(16, 20, 1234), dtype="complex64")
    assert prog.stats["tile_bits"] == 11
    for step in prog.passes[:2]:
        rc, size, log = _build(step, L.QSV_C64)
        if rc == L.QSV_EIO:
            pytest.skip(f"NVRTC not installed: {log}")
        assert rc == 0 and size > 10000, log


def test_other_tile_sizes_are_left_to_the_interpreter():
    prog = compile_circuit(W.random_1q_cz(12, 6, 1), tile_bits=8)
    rc, size, log = _build(prog.passes[0])
    assert rc == L.QSV_EINVAL and "not eligible" in log


# ------------------------------------------------------------------ scatter passes (pass fused with a swap)
import random
import re


def _source(step, scatter_bits=None, dtype=L.QSV_C128):
    buf, need = C.create_string_buffer(1 << 20), C.c_size_t()
    lib = L.load()
    if scatter_bits is None:
        rc = lib.qsv_jit_source(C.byref(step.desc), step.ops, dtype, buf, len(buf), C.byref(need))
    else:
        lb = (C.c_int * len(scatter_bits))(*scatter_bits)
        rc = lib.qsv_jit_source_scatter(C.byref(step.desc), step.ops, dtype, len(scatter_bits), lb, buf, len(buf), C.byref(need))
    assert rc == 0 and need.value < len(buf)
    return buf.value.decode()


def _plain_stores(src):
    """(gt, base) -> [final local index of register j]: the store addressing of the plain kernel."""
    thr = [(int(a), int(b)) for a, b in re.findall(r"gb \|= \(unsigned long long\)\(\(gt >> (\d+)\) & 1u\) << (\d+);", src)]
    st = [(int(a), int(f), int(j)) for a, f, j in re.findall(r"state\[\(gb \| (\d+)ull\) \^ (\d+)ull\] = v\[(\d+)\];", src)]
    assert len(thr) == 7 and len(st) == 16

    def f(gt, base):
        gb = base
        for i, p in thr:
            gb |= ((gt >> i) & 1) << p
        return [(gb | a) ^ flip for a, flip, _ in sorted(st, key=lambda t: t[2])]
    return f


def _scatter_stores(src):
    """(gt, base, keep) -> [(target buffer, index there) of register j]: the scatter kernel's addressing."""
    thr = [(int(a), int(b)) for a, b in re.findall(r"gb \|= \(unsigned long long\)\(\(gt >> (\d+)\) & 1u\) << (\d+);", src)]
    gxk = int(re.search(r"const unsigned long long gx = gb \^ (\d+)ull;", src).group(1))
    sels = [(int(a), int(b)) for a, b in re.findall(r"sel \|= \(unsigned\)\(\(gx >> (\d+)\) & 1ull\) << (\d+);", src)]
    lmask = int(re.search(r"gk = \(gx & ~(\d+)ull\) \| D\.keep;", src).group(1))
    one = [(0, int(a), int(j)) for a, j in re.findall(r"dst\[gk \| (\d+)ull\] = v\[(\d+)\];", src)]
    many = [(int(s), int(a), int(j)) for s, a, j in re.findall(r"D\.p\[sel \| (\d+)u\]\[gk \| (\d+)ull\] = v\[(\d+)\];", src)]
    st = one or many
    assert len(thr) == 7 and len(st) == 16 and not (one and many)
    assert ("JV *const dst = D.p[sel];" in src) == bool(one)

    def f(gt, base, keep):
        gb = base
        for i, p in thr:
            gb |= ((gt >> i) & 1) << p
        gx = gb ^ gxk
        sel = 0
        for l, i in sels:
            sel |= ((gx >> l) & 1) << i
        gk = (gx & ~lmask) | keep
        return [(sel | sj, gk | a) for sj, a, _ in sorted(st, key=lambda t: t[2])]
    return f, lmask


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_scatter_kernel_addressing_is_the_plain_store_followed_by_the_swap(dtype):
    """The scatter variant of a pass must send the amplitude the plain kernel stores at local index ix
    to buffer x = (bits of ix at the swapped local positions) at index ix with those bits replaced by
    this rank's bits — for swapped positions held in registers, on thread bits, and outside the tile."""
    n = 18
    prog = compile_circuit(W.random_1q_cz(n, 12, 7), dtype=dtype)
    rng = random.Random(5)
    code = L.QSV_C128 if dtype == "complex128" else L.QSV_C64
    checked = set()
    for step in prog.passes:
        d = step.desc
        last = d.rounds[d.n_rounds - 1]
        store = [d.store_bits[i] for i in range(d.n_tile)]
        regs = [store[last.reg_pos[q]] for q in range(4)]
        thrs = [b for b in store if b not in regs]
        outside = [b for b in range(n) if b not in store]
        plain = _plain_stores(_source(step, None, code))
        for name, bits in (("outside", outside[:2]), ("thread", thrs[1:3]), ("register", [regs[0], regs[3]]),
                           ("mixed", [regs[1], thrs[0], outside[-1]])):
            scat, lmask = _scatter_stores(_source(step, bits, code))
            assert lmask == sum(1 << b for b in bits)
            for _ in range(40):
                base = rng.getrandbits(n) & ~sum(1 << b for b in store)
                gt = rng.randrange(128)
                keep = sum(rng.getrandbits(1) << b for b in bits)
                for ix, (x, idx) in zip(plain(gt, base), scat(gt, base, keep)):
                    assert x == sum(((ix >> b) & 1) << i for i, b in enumerate(bits)), (name, bits)
                    assert idx == (ix & ~lmask) | keep, (name, bits)
            checked.add(name)
    assert checked == {"outside", "thread", "register", "mixed"}


def test_scatter_kernels_compile_for_sm_100a():
    prog = compile_circuit(W.random_1q_cz(18, 12, 7))
    step = prog.passes[-1]
    d = step.desc
    store = [d.store_bits[i] for i in range(d.n_tile)]
    outside = [b for b in range(18) if b not in store]
    for bits in ([outside[-1]], [store[10], outside[0]], [store[3], store[9], outside[1]]):
        n, log = C.c_size_t(), C.create_string_buffer(1 << 14)
        lb = (C.c_int * len(bits))(*bits)
        rc = L.load().qsv_jit_build_scatter(C.byref(d), step.ops, L.QSV_C128, len(bits), lb, C.byref(n), log, len(log))
        if rc == L.QSV_EIO:
            pytest.skip("NVRTC not installed")
        assert rc == 0 and n.value > 10000, log.value.decode(errors="replace")


def test_scatter_needs_full_tile_coverage():
    """Zero-support skipping leaves tiles unvisited, which an out-of-place pass cannot allow."""
    prog = compile_circuit(W.ghz(16), skip_zero_support=True)
    skipping = [s for s in prog.passes if s.desc.n_active >= 0]
    assert skipping
    lb = (C.c_int * 1)(15)
    need = C.c_size_t()
    rc = L.load().qsv_jit_source_scatter(C.byref(skipping[0].desc), skipping[0].ops, L.QSV_C128, 1, lb, None, 0, C.byref(need))
    assert rc == L.QSV_EINVAL
