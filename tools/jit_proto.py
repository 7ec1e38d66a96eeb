"""Prototype: emit a specialised (straight-line) ring kernel for one compiled pass; compile; count."""
import sys, time, subprocess, re, collections
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from quantum_simulations_b200 import _lib as L, workloads as W
from quantum_simulations_b200.kernel.cuda_dense import compile_circuit

T, RB = 11, 4

def swz(x):
    f = x
    for s in range(3, 16, 3):
        f ^= (x >> s) & 7
    return f

def gen(step, name="k_pass_jit"):
    d = step.desc
    load = [d.load_bits[i] for i in range(T)]
    store = [d.store_bits[i] for i in range(T)]
    out = []
    A = out.append
    A('#include "jit_prelude.cuh"')
    A(f'extern "C" __global__ void __launch_bounds__(512, 1) {name}(double2 *__restrict__ state, const double2 *__restrict__ tables, const unsigned long long rank_bits, const unsigned n_tiles) {{')
    A('  JIT_RING_PROLOGUE')
    # tile base insertion
    A('  #define TILE_BASE(tile) ({ unsigned long long b_ = (tile); ' + ' '.join(f'b_ = insert_zero_bit(b_, {p});' for p in load) + ' b_; })')
    # producer
    A('  if (tid >= 384) {')
    A('    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");')
    A('    const unsigned pt = tid - 384;')
    A('    unsigned long long off_lo = 0;')
    for i in range(7):
        A(f'    off_lo |= (unsigned long long)((pt >> {i}) & 1u) << {load[i]};')
    A('    const unsigned sw_lo = tile_swizzle3(pt);')
    A('    unsigned s = 0;')
    A('    for (unsigned tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++s) {')
    A('      const int b = s % 6; const unsigned use = s / 6;')
    A('      const double2 *g = state + TILE_BASE(tile) + off_lo;')
    A('      double2 *dd = S.buf[b];')
    A('      if (use > 0) mbar_wait_sleep(&S.empty[b], (use - 1) & 1);')
    for k in range(16):
        o = sum(1 << load[7 + i] for i in range(4) if k & (1 << i))
        A(f'      cp_async16(dd + (sw_lo ^ {swz(k << 7)}u), g + {o}ull);')
    A('      cp_async_arrive(&S.full[b]);')
    A('    }')
    A('    return;')
    A('  }')
    A('  asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");')
    A('  const int grp = tid >> 7; const unsigned gt = tid & 127u;')
    A('  unsigned s = grp;')
    A('  for (unsigned tile = blockIdx.x + (unsigned)grp * gridDim.x; tile < n_tiles; tile += gridDim.x * 3, s += 3) {')
    A('    const int b = s % 6; const unsigned use = s / 6;')
    A('    double2 *buf = S.buf[b];')
    A('    const unsigned long long base = TILE_BASE(tile);')
    A('    const unsigned long long glob = rank_bits | base;')
    A('    mbar_wait(&S.full[b], use & 1);')
    A('    double2 v[16];')
    nr = d.n_rounds
    for r in range(nr):
        rd = d.rounds[r]
        regs = [rd.reg_pos[q] for q in range(RB)]
        thr = [rd.thr_pos[i] for i in range(T - RB)]
        A(f'    {{ // round {r}')
        A('      unsigned xb = 0;')
        for i, p in enumerate(thr):
            A(f'      xb |= ((gt >> {i}) & 1u) << {p};')
        A('      const unsigned sb = tile_swizzle3(xb);')
        if rd.fold_off >= 0:
            A(f'      const double2 fold = __ldg(&tables[{rd.fold_off} + gt]);')
        for j in range(16):
            c = 0
            for q in range(RB):
                if j & (1 << q):
                    c ^= swz(1 << regs[q])
            A(f'      v[{j}] = buf[sb ^ {c}u];')
        last = r == nr - 1
        if last:
            A('      mbar_arrive(&S.empty[b]);')
        if rd.fold_off >= 0:
            A('      apply_fold(v, fold);')
        for o in range(rd.op_begin, rd.op_end):
            op = step.ops[o]
            m = [repr(float(op.m[k])) for k in range(4)]
            k_ = op.kind
            tb = op.target
            if k_ in (L.OP_HAD, L.OP_ROT) and (op.flags or not (op.reg_ctrl or op.tile_ctrl or op.glob_ctrl)):
                if op.flags:
                    A(f'      {{ const int sm = (int)(((unsigned)__popc(xb & {op.tile_ctrl}u) + (unsigned)__popcll(glob & {op.glob_ctrl}ull) + {op.flags >> 1}u) << 31);')
                    A(f'        op_pre<double2, double, {tb}>(v, {op.flags}u, sm, {m[2]}, {m[3]}); }}')
                if k_ == L.OP_HAD:
                    A(f'      op_had<double2, {tb}>(v);')
                else:
                    A(f'      op_rot<double2, double, {tb}, false>(v, {m[0]}, {m[1]}, 0u);')
                continue
            cond = []
            if op.glob_ctrl:
                cond.append(f'(glob & {op.glob_ctrl}ull) == {op.glob_ctrl}ull')
            if op.tile_ctrl:
                cond.append(f'(xb & {op.tile_ctrl}u) == {op.tile_ctrl}u')
            pre = f'if ({" && ".join(cond)}) ' if cond else ''
            rc = op.reg_ctrl
            if k_ == L.OP_ROT:
                A(f'      {pre}op_rot<double2, double, {tb}, true>(v, {m[0]}, {m[1]}, {rc}u);')
            elif k_ == L.OP_XSWAP:
                A(f'      {pre}op_xswap<double2, {tb}, true>(v, {rc}u);')
            elif k_ == L.OP_YSWAP:
                A(f'      {pre}op_yswap<double2, {tb}, true>(v, {rc}u);')
            elif k_ == L.OP_PHASE:
                A(f'      {pre}op_phase_mask<double2, double>(v, {m[0]}, {m[1]}, {rc}u);')
            elif k_ == L.OP_SIGN:
                A(f'      {pre}op_sign_mask<double2>(v, {rc}u);')
            elif k_ == L.OP_SCALE:
                A(f'      op_scale<double2, double>(v, {m[0]});')
            else:
                raise ValueError(k_)
        if last:
            A('      unsigned long long gb = base;')
            for i, p in enumerate(thr):
                A(f'      gb |= (unsigned long long)((gt >> {i}) & 1u) << {store[p]};')
            for j in range(16):
                a = 0
                for q in range(RB):
                    if j & (1 << q):
                        a |= 1 << store[regs[q]]
                A(f'      state[(gb | {a}ull) ^ {int(d.store_flip)}ull] = v[{j}];')
        else:
            for j in range(16):
                c = 0
                for q in range(RB):
                    if j & (1 << q):
                        c ^= swz(1 << regs[q])
                A(f'      buf[sb ^ {c}u] = v[{j}];')
            A('      asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(128) : "memory");')
        A('    }')
    A('  }')
    A('}')
    return '\n'.join(out)

if __name__ == '__main__':
    n = 30
    prog = compile_circuit(W.random_1q_cz(n, 20, 1234), low_bits=3)
    step = max(prog.passes, key=lambda s: s.n_micro_ops)
    src = gen(step)
    Path('/tmp/jit/k.cu').write_text(src)
    print('ops', step.n_micro_ops, 'rounds', step.desc.n_rounds, 'lines', src.count('\n'))
