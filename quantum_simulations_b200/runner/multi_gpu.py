"""One process per GPU: execute a sharded program (circuit/sharding.py) on this rank's shard.

The reference's multi-chunk runner walks chunk files and, for a gate on a "non-local" qubit,
loads the 2 or 4 chunks of a group together (wenbo_engine/runner/single_node.py:219-321).
Here every rank keeps its 2^n_local amplitudes in HBM for the whole run; stages only mix local
qubits (zero communication, rank bits enter as constants) and are connected by
``SwapStep``s = one NCCL all-to-all of the swapped blocks over NVLink (csrc/exchange.cuh).

The host side needs only a rendezvous and a few tiny collectives (NCCL unique id, CUDA IPC handles,
"do all ranks agree", sums of marginals): runner/plumbing.py does that over plain TCP (stdlib);
no amplitude ever passes through it, and the product imports no tensor library.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

from quantum_simulations_b200 import _lib as L
from quantum_simulations_b200.circuit import sharding
from quantum_simulations_b200.circuit.io import validate_circuit_dict
from quantum_simulations_b200.circuit.passes import PassStep, Program, SwapStep


def execute(prog: Program, backend) -> None:
    """Run `prog` on one shard.  `backend` provides run_passes(steps, first, count) over a maximal run of
    consecutive passes and swap(global_bits, local_bits).  A backend that also has
    ``transition(swap_step)`` / ``swap_pipelined(...)`` (CudaShard after prepare()) executes the swaps it
    has a plan for PIPELINED with the passes next to them (sharding.plan_transitions); the passes a
    transition consumed are skipped here."""
    steps = list(prog.steps)
    runs: list = []                      # (first index, [PassStep...]) maximal runs
    k = 0
    while k < len(steps):
        if isinstance(steps[k], PassStep):
            j = k
            while j < len(steps) and isinstance(steps[j], PassStep):
                j += 1
            runs.append((k, steps[k:j]))
            k = j
        else:
            if not isinstance(steps[k], SwapStep):
                raise TypeError(f"sharded programs hold passes and swaps only, got {type(steps[k]).__name__}")
            k += 1
    run_at = {first: run for first, run in runs}
    run_end = {first + len(run): (first, run) for first, run in runs}
    plan_of = getattr(backend, "transition", lambda _s: None)
    done_head: dict = {}                 # first index of a run -> passes already executed by a transition
    k = 0
    while k < len(steps):
        st = steps[k]
        if isinstance(st, PassStep):
            run = run_at[k]
            head = done_head.get(k, 0)
            nxt = k + len(run)
            tr = plan_of(steps[nxt]) if nxt < len(steps) else None
            tail = tr.a_count if tr is not None else 0
            scatter = tr is None and nxt < len(steps) and getattr(backend, "fused_exchange", False) and head < len(run)
            if scatter:
                tail = 1                 # scatter pass (opt-in): the last pass of the run is fused with the exchange
            if len(run) - head - tail > 0:
                if head == 0 and tail == 0:
                    backend.run_passes(run)
                else:
                    backend.run_passes(run, head, len(run) - head - tail)
            if scatter:
                backend.scatter_swap(run, list(steps[nxt].global_bits), list(steps[nxt].local_bits))
                nxt += 1
            k = nxt
            continue
        tr = plan_of(st)
        if tr is None:
            backend.swap(list(st.global_bits), list(st.local_bits))
        else:
            before = run_end.get(k)
            after = run_at.get(k + 1)
            backend.swap_pipelined(before[1] if before else None, tr.a_count, after if after else None, tr.b_count,
                                   list(st.global_bits), list(st.local_bits), list(tr.chunk_bits))
            if after is not None:
                done_head[k + 1] = tr.b_count
        k += 1


class CudaShard:
    """backend of `execute` over libqsv: a DeviceState for (rank, world) with an NCCL communicator."""

    def __init__(self, n_qubits: int, rank: int, world: int, dtype="complex128", device: int | None = None,
                 unique_id: bytes | None = None, local: bool = False):
        """local=True: this shard is one of several handles of ONE process (wire them with
        ``CudaShard.wire_local``): no NCCL communicator, swaps run on the flag-ordered exchange kernels."""
        from quantum_simulations_b200.kernel.cuda import DeviceState
        self.state = DeviceState(n_qubits, dtype, rank if device is None else device, rank, world)
        self.rank, self.world = rank, world
        self._uploaded: dict = {}
        self.peer_swap, self.peer_error = False, None
        self.fused_exchange, self.fused_error = False, None
        self.swaps = self.pipelined_swaps = self.fused_swaps = 0
        self._transitions: dict = {}          # id(SwapStep) -> (SwapStep, Transition)
        self.pipeline = os.environ.get("QSV_PIPELINE", "1") != "0"
        self.xchg_sms = int(os.environ.get("QSV_XCHG_SMS", "20"))
        if world > 1 and local:
            self.state._ck(self.state.lib.qsv_comm_init_local(self.state._h))
        elif world > 1:
            if unique_id is None or len(unique_id) != 128:
                raise ValueError("world > 1 needs the 128-byte NCCL unique id of rank 0 (nccl_unique_id())")
            buf = C.create_string_buffer(unique_id, 128)
            self.state._ck(self.state.lib.qsv_comm_init(self.state._h, buf))

    @staticmethod
    def wire_local(shards) -> None:
        """Shards of one process: hand every handle the device pointers of all of them."""
        world = len(shards)
        ptrs = (C.c_void_p * world)(*[s.state.device_ptr()[0] for s in shards])
        for s in shards:
            s.state._ck(s.state.lib.qsv_comm_set_peers_local(s.state._h, ptrs))
            s.peer_swap = True

    def map_peers(self, dist) -> bool:
        """Exchange CUDA IPC handles of the shards (plumbing: an all-gather of 64 bytes per rank) so
        that swaps run as one peer-memory kernel over NVLink.  Returns False (NCCL send/recv swaps
        stay in use) if the box does not allow IPC mappings; QSV_SWAP=nccl skips the attempt."""
        self.peer_error = None
        if self.world == 1 or os.environ.get("QSV_SWAP", "peer") == "nccl":
            return False
        lib, h = self.state.lib, self.state._h
        mine = C.create_string_buffer(64)
        ok = lib.qsv_comm_ipc_handle(h, mine) == 0
        box = dist.all_gather_object(mine.raw if ok else None)
        if any(b is None for b in box):
            self.peer_error = "cudaIpcGetMemHandle failed on some rank"
            mapped = False
        else:
            mapped = lib.qsv_comm_set_peers(h, C.create_string_buffer(b"".join(box), 64 * self.world)) == 0
            if not mapped:
                self.peer_error = (lib.qsv_last_error(h) or b"?").decode(errors="replace")
        self.peer_swap = dist.all(bool(mapped))
        lib.qsv_set_option(h, L.OPT_PEER_SWAP, int(self.peer_swap))     # all ranks must take the same path
        return self.peer_swap

    def map_shadows(self, dist) -> bool:
        """Opt-in (after map_peers): give every shard a second buffer and map the peers' second buffers
        too, so that a pass followed by a swap runs as ONE kernel that stores its results where they live
        after the swap (qsv_pass_scatter).  Costs 2x the shard in HBM; returns False, and changes
        nothing, if any rank cannot allocate or map."""
        self.fused_error = None
        if self.world == 1 or not self.peer_swap:
            return False
        lib, h = self.state.lib, self.state._h
        mine = C.create_string_buffer(64)
        ok = lib.qsv_comm_shadow_ipc_handle(h, mine) == 0
        if not ok:
            self.fused_error = (lib.qsv_last_error(h) or b"?").decode(errors="replace")
        box = dist.all_gather_object(mine.raw if ok else None)
        mapped = False
        if all(b is not None for b in box):
            mapped = lib.qsv_comm_set_shadow_peers(h, C.create_string_buffer(b"".join(box), 64 * self.world)) == 0
            if not mapped:
                self.fused_error = (lib.qsv_last_error(h) or b"?").decode(errors="replace")
        self.fused_exchange = dist.all(bool(mapped))     # all ranks must take the same path
        return self.fused_exchange

    def prepare(self, prog: Program, agree=None) -> None:
        """Upload (and specialise) every run of passes once; execute() then only replays.  Swaps get a
        pipelined transition (sharding.plan_transitions) when the peers are mapped and every pass it names
        is specialised ON EVERY RANK: `agree(flag) -> bool` is the collective AND over the ranks
        (ShardedSimulator passes an all-reduce; shards of one process are agreed by their driver)."""
        steps = list(prog.steps)
        handle_at: dict = {}
        run: list = []
        for k, step in enumerate(steps + [None]):
            if isinstance(step, PassStep):
                run.append(step)
            elif run:
                h = self.state.upload_steps(run)
                self._uploaded[id(run[0])] = (run[0], h)       # the step object is kept alive: its id cannot be recycled
                for i in range(len(run)):
                    handle_at[k - len(run) + i] = (h, i)
                if self.fused_exchange and isinstance(step, SwapStep):
                    l = (C.c_int * len(step.local_bits))(*step.local_bits)
                    ok = self.state.lib.qsv_pass_scatter_prepare(self.state._h, h, len(run) - 1, len(step.local_bits), l) == 0
                    # the library falls back to "pass, then swap" when it has no scatter kernel: every rank must
                    # take the same branch, so one rank without the kernel switches the scatter path off everywhere
                    if agree is not None and not agree(bool(ok)):
                        self.fused_exchange = False
                run = []
        if not (self.pipeline and self.peer_swap and not self.fused_exchange):
            return
        plans = sharding.plan_transitions(prog, min_chunk_pos=8 if prog.n_local >= 24 else 5)   # >= 4 KB runs = one exchange unit
        for k in sorted(plans):
            tr = plans[k]
            ok = True
            for idx in [k - 1 - i for i in range(tr.a_count)] + [k + 1 + i for i in range(tr.b_count)]:
                h, i = handle_at[idx]
                n = 64
                flags = (C.c_int * n)()
                # programs hold at most a few dozen passes; a longer one is simply not pipelined
                ok = ok and i < n and self.state.lib.qsv_program_specialised(self.state._h, h, flags, n) == 0 and bool(flags[i])
            if agree is not None:
                ok = agree(bool(ok))
            if ok:
                self._transitions[id(steps[k])] = (steps[k], tr)

    def transition(self, swap_step):
        e = self._transitions.get(id(swap_step))
        return e[1] if e is not None and e[0] is swap_step else None

    def _handle_of(self, steps):
        e = self._uploaded.get(id(steps[0]))
        return e[1] if e is not None and e[0] is steps[0] else None

    def release(self, prog: Program | None = None) -> None:
        """Forget the device programs (and transitions) of `prog` (default: of every prepared program)."""
        keep = {}
        mine = None if prog is None else {id(s) for s in prog.steps}
        for k, (step, h) in self._uploaded.items():
            if mine is None or k in mine:
                self.state.release_program(h)
            else:
                keep[k] = (step, h)
        self._uploaded = keep
        self._transitions = {k: v for k, v in self._transitions.items() if mine is not None and k not in mine}

    def run_passes(self, steps, first: int = 0, count: int | None = None) -> None:
        count = len(steps) - first if count is None else count
        h = self._handle_of(steps)
        temp = h is None
        if temp:
            h = self.state.upload_steps(steps)
        self.state._ck(self.state.lib.qsv_program_run_range(self.state._h, h, first, count))
        if temp:
            self.state.release_program(h)

    def swap_pipelined(self, before, a_count, after, b_count, global_bits, local_bits, chunk_bits) -> None:
        """The swap together with the last a_count passes of `before` and the first b_count passes of
        `after` (lists of PassStep prepared on this shard), chunk by chunk (qsv_swap_pipelined)."""
        st, lib = self.state, self.state.lib
        ha = self._handle_of(before) if a_count else None
        hb = self._handle_of(after) if b_count else None
        if (a_count and ha is None) or (b_count and hb is None):
            raise RuntimeError("swap_pipelined: the neighbouring passes were not prepared on this shard")
        s = len(global_bits)
        g, l = (C.c_int * s)(*global_bits), (C.c_int * s)(*local_bits)
        cb = (C.c_int * len(chunk_bits))(*chunk_bits)
        st._ck(lib.qsv_swap_pipelined(st._h, ha, (len(before) - a_count) if a_count else 0, a_count, hb, b_count,
                                      s, g, l, len(chunk_bits), cb, self.xchg_sms))
        self.swaps += 1
        self.pipelined_swaps += 1

    def scatter_swap(self, steps, global_bits, local_bits) -> None:
        """Scatter pass (opt-in, 2x memory): the LAST pass of `steps` fused with the exchange (qsv_pass_scatter)."""
        h = self._handle_of(steps)
        temp = h is None
        if temp:
            h = self.state.upload_steps(steps)
        st, lib = self.state, self.state.lib
        s = len(global_bits)
        g = (C.c_int * s)(*global_bits)
        l = (C.c_int * s)(*local_bits)
        ov = C.c_int(0)
        st._ck(lib.qsv_pass_scatter(st._h, h, len(steps) - 1, s, g, l, C.byref(ov)))
        self.fused_swaps += int(ov.value)
        self.swaps += 1
        if temp:
            self.state.release_program(h)

    def swap(self, global_bits, local_bits) -> None:
        self.swaps += 1
        s = len(global_bits)
        g = (C.c_int * s)(*global_bits)
        l = (C.c_int * s)(*local_bits)
        self.state._ck(self.state.lib.qsv_swap_global_local(self.state._h, s, g, l))

    def close(self) -> None:
        self._uploaded = {}                   # the handle owns the device programs
        self.state.close()


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = L.load().qsv_comm_unique_id(buf)
    if rc:
        raise L.QsvError(rc, "ncclGetUniqueId failed (is libnccl.so.2 loadable?)")
    return buf.raw


# ------------------------------------------------------------------ host plumbing (runner/plumbing.py)
def dist_env() -> tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_plumbing():
    """The process-wide HostPlumbing over MASTER_ADDR / MASTER_PORT (set by torchrun)."""
    from quantum_simulations_b200.runner.plumbing import init_plumbing as _init
    return _init()


def share_unique_id(dist, rank: int) -> bytes:
    return dist.broadcast_object(nccl_unique_id() if rank == 0 else None, src=0)


class ShardedSimulator:
    """Public multi-GPU surface (run under torchrun, one process per GPU).  The constructor does
    the one-time work (rendezvous, NCCL communicator, shard allocation); ``simulate`` runs one
    circuit from |0...0> and returns THIS rank's shard in host memory."""

    def __init__(self, n_qubits: int, dtype="complex128", fused_exchange: bool | None = None):
        """fused_exchange (default: env QSV_FUSED_EXCHANGE=1): second buffer per shard, passes before a
        swap run as scatter passes (see CudaShard.map_shadows)."""
        self.rank, self.local_rank, self.world = dist_env()
        self.n = n_qubits
        self.g = int(math.log2(self.world))
        if 1 << self.g != self.world:
            raise ValueError("world size must be a power of two")
        self.dtype = np.dtype(dtype)
        self.dist = init_plumbing() if self.world > 1 else None
        uid = share_unique_id(self.dist, self.rank) if self.world > 1 else None
        self.shard = CudaShard(n_qubits, self.rank, self.world, dtype, self.local_rank, uid)
        self.peer_swap = self.shard.map_peers(self.dist) if self.world > 1 else False
        if fused_exchange is None:
            fused_exchange = os.environ.get("QSV_FUSED_EXCHANGE", "0") == "1"
        self.fused_exchange = bool(fused_exchange) and self.peer_swap and self.shard.map_shadows(self.dist)
        self.logical_rank, self._flip_mask = self.rank, 0
        self._prepared: dict = {}

    def plan(self, circuit_dict: dict, **compiler_kw) -> Program:
        from quantum_simulations_b200.kernel.cuda_dense import circuit_ops
        cd = validate_circuit_dict(circuit_dict)
        if cd["number_of_qubits"] != self.n:
            raise ValueError("circuit size differs from the simulator's")
        return self.plan_ops(circuit_ops(cd), **compiler_kw)

    def plan_ops(self, ops: list, **compiler_kw) -> Program:
        """Stage plan for a step-IR op list [(qubits, U)] (what the front ends produce: circuit dicts,
        circuit/qasm.qasm_to_ops, circuit/hisvsim_parts)."""
        compiler_kw.setdefault("swap_anywhere", bool(self.peer_swap))     # peer kernel: no relabel before a swap
        compiler_kw.setdefault("rank_flips", True)                        # X pending on a rank bit renames shards
        if getattr(self, "fused_exchange", False):
            compiler_kw.setdefault("fused_exchange", True)                # cost model: a swap hides the pass before it
        return sharding.plan(ops, self.n, self.n - self.g, self.dtype.name, **compiler_kw)

    def plan_parts(self, parts_ops: list, **compiler_kw) -> Program:
        """Stage plan that follows a given PARTITION of the circuit (HiSVSIM part files,
        circuit/hisvsim_parts.qasm_parts): one stage per part, one gather of the qubits it mixes in front of it
        (sharding.plan_parts)."""
        compiler_kw.setdefault("swap_anywhere", bool(self.peer_swap))
        return sharding.plan_parts(parts_ops, self.n, self.n - self.g, self.dtype.name, **compiler_kw)

    def simulate_qasm(self, text: str, out: np.ndarray | None = None, **compiler_kw) -> np.ndarray:
        """OpenQASM 2.0 program from |0...0> on the sharded state (front end as in
        kernel.cuda_dense.simulate_qasm); returns this rank's logical shard like ``simulate``."""
        from quantum_simulations_b200.circuit.fusion import fuse_2q_blocks
        from quantum_simulations_b200.circuit.qasm import qasm_to_ops
        n, ops = qasm_to_ops(text)
        if n != self.n:
            raise ValueError("circuit size differs from the simulator's")
        self.run(self.plan_ops(fuse_2q_blocks(ops, tol=1e-14), **compiler_kw))
        return self.shard.state.download(out)

    def simulate(self, circuit_dict: dict, out: np.ndarray | None = None, sink=None, chunk_amps: int = 1 << 24,
                 **compiler_kw):
        """Returns the amplitudes of LOGICAL shard ``self.logical_rank`` (= rank ^ the program's
        rank_flip_mask: an X gate left pending on a rank bit is a renaming of the shards, not a
        data movement); index of amplitude i of the result = (logical_rank << n_local) | i.
        With ``sink(view, first_amp)`` the shard is not returned but streamed to the sink through two pinned
        staging buffers of chunk_amps amplitudes (states larger than host memory: 36 qubits = 128 GiB per GPU)."""
        prog = self.plan(circuit_dict, **compiler_kw)
        self.run(prog)
        if sink is not None:
            return self.shard.state.stream_to_host(sink, chunk_amps)
        return self.shard.state.download(out)

    def _agree(self, flag: bool) -> bool:
        """Collective AND over the ranks (host plumbing): every rank takes the same execution path."""
        return bool(flag) if self.dist is None else self.dist.all(flag)

    def prepare(self, prog: Program) -> None:
        """Upload and specialise the passes of `prog` and plan its pipelined transitions (collective)."""
        self.shard.prepare(prog, self._agree)
        self._prepared[id(prog)] = prog

    def run(self, prog: Program, init: bool = True) -> None:
        """init=False: the program continues from what the shards hold (a resumed / segmented run)."""
        temp = self._prepared.get(id(prog)) is not prog
        if temp:
            self.shard.prepare(prog, self._agree)
        if init and not prog.fused_init:       # otherwise the first pass creates |0...0> itself
            self.shard.state.init_zero()
        execute(prog, self.shard)
        if temp:
            self.shard.release(prog)
        self.logical_rank = self.rank ^ prog.rank_flip_mask
        self._flip_mask = prog.rank_flip_mask

    def probabilities(self, qubits) -> np.ndarray:
        """Marginal distribution over logical `qubits` of the whole sharded state (identical on every
        rank): per-shard marginals (qsv_probabilities, rank bits allowed) summed over the shards."""
        p = self.shard.state.probabilities(self._physical(qubits))
        if self.world > 1:
            p = self.dist.allreduce(p, "sum")
        return p

    def expect_z(self, qubits) -> float:
        v = float(self.shard.state.expect_z(self._physical(qubits)))
        return self.dist.allreduce(v, "sum") if self.world > 1 else v

    def _physical(self, qubits) -> list:
        """After a run the layout is the identity except for the shard renaming of rank_flip_mask: a
        flipped rank bit is seen inverted by the shard-local kernels, which read the PHYSICAL rank."""
        n_loc = self.n - self.g
        if any(q >= n_loc and (self._flip_mask >> (q - n_loc)) & 1 for q in qubits):
            raise NotImplementedError("observable on a qubit whose rank bit is renamed (rank_flip_mask): "
                                      "use sample() or gather the shards")
        return list(qubits)

    def sample(self, seed: int, shots: int) -> np.ndarray:
        """Measurement samples of the sharded state, identical on every rank and bit-exact with
        oracle/ref_dense.py::sample_indices.  The exclusive scan over leaf sums is one sequential
        chain over all shards (rank order = index order), evaluated redundantly on every rank
        from the all-gathered leaf sums (8 bytes per 1024 amplitudes); every rank then walks the
        leaves of the shots that fall into its shard."""
        st, lib = self.shard.state, self.shard.state.lib
        n_loc = self.n - self.g
        leaf_log2 = min(10, n_loc)
        n_leaves = (1 << n_loc) >> leaf_log2
        mine = np.empty(n_leaves, dtype=np.float64)
        st._ck(lib.qsv_leaf_sums(st._h, mine.ctypes.data_as(C.POINTER(C.c_double))))
        if self.world > 1:
            parts = self.dist.all_gather_object(mine)
            sums = np.concatenate([parts[l ^ self._flip_mask] for l in range(self.world)])   # logical order
        else:
            sums = mine
        offs = np.empty(len(sums) + 1)
        offs[0] = 0.0
        np.cumsum(sums, out=offs[1:])                 # np.cumsum is a sequential left-to-right scan
        total = offs[-1]
        x = np.sort(np.random.default_rng(seed).random(shots)) * total
        b = np.searchsorted(offs[1:], x, side="right")
        out = np.zeros(shots, dtype=np.int64)
        over = b >= len(sums)
        out[over] = (1 << self.n) - 1
        lo, hi = self.logical_rank * n_leaves, (self.logical_rank + 1) * n_leaves
        sel = np.nonzero((b >= lo) & (b < hi) & ~over)[0]
        if len(sel):
            lidx = np.ascontiguousarray(b[sel] - lo, dtype=np.uint64)
            loff = np.ascontiguousarray(offs[b[sel]])
            xs = np.ascontiguousarray(x[sel])
            got = np.empty(len(sel), dtype=np.uint64)
            st._ck(lib.qsv_sample_in_leaves(st._h, len(sel), lidx.ctypes.data_as(C.POINTER(C.c_uint64)),
                                            loff.ctypes.data_as(C.POINTER(C.c_double)),
                                            xs.ctypes.data_as(C.POINTER(C.c_double)),
                                            got.ctypes.data_as(C.POINTER(C.c_uint64))))
            out[sel] = got.astype(np.int64) + (self.logical_rank << n_loc)
        if self.world > 1:
            if self.rank != 0:
                out[over] = 0                         # the clamp is contributed once (rank 0)
            parts = self.dist.all_gather_object(out)  # integer sum (indices exceed float64's exact range at n > 53)
            out = np.sum(np.stack(parts), axis=0, dtype=np.int64)
        return out.astype(np.uint64)

    def close(self) -> None:
        self.shard.close()


def simulate_sharded(circuit_dict: dict, dtype="complex128", out: np.ndarray | None = None, **compiler_kw):
    """One-shot form of ShardedSimulator."""
    cd = validate_circuit_dict(circuit_dict)
    sim = ShardedSimulator(cd["number_of_qubits"], dtype)
    try:
        return sim.simulate(cd, out, **compiler_kw)
    finally:
        sim.close()


def run(circuit_dict: dict, work_dir, chunk_size: int = 1 << 20, dtype: str = "complex128",
        use_wal: bool = True, checkpoint_every: int = 0, stop_after_checkpoints: int | None = None,
        use_staging: bool = False, staging_method: str = "heuristic", **compiler_kw):
    """Multi-GPU form of ``runner.single_node.run`` (reference single_node.py:78-176), launched with one process per
    GPU.  The circuit is cut into SEGMENTS of `checkpoint_every` levels (0 = one segment); each segment is planned
    and executed as a sharded program that starts and ends in the identity layout, and after each segment every
    rank writes the chunk files of its shard through two pinned staging buffers; rank 0 publishes the manifest and
    commits the WAL (done_steps = levels completed — the same count the single-device runner and the reference
    commit, so a work directory is resumable by either) once all chunks are durable.  A restart loads the
    committed buffer into the shards and continues with the next segment (reference resume:
    single_node.py:143-176, wal/wal.py:85-89); a directory whose WAL already covers the circuit returns at once.
    stop_after_checkpoints (tests): return after that many checkpoints, as a crash would.
    use_staging=True (reference single_node.py:108-134): the stage structure comes from the reference's
    ``atlas_stages`` (staging_method "heuristic" | "greedy" | "ilp") instead of the engine's own planner
    (sharding.plan_atlas: each atlas stage = one stage here); the run is then one segment, the chunks are written
    in atlas's PHYSICAL layout and ``qubit_mapping.json`` holds log_to_phys, exactly as the reference does, so
    ``collect_state(buf, apply_permutation=True, work_dir=...)`` returns the logical state.
    Returns the committed buffer path (``collect_state`` reads it exactly like a single-device run)."""
    from pathlib import Path

    from quantum_simulations_b200.runner.single_node import _buf_dir, _other, _wipe_buf, build_steps
    from quantum_simulations_b200.storage.block_store import chunk_filename, read_chunk, write_chunk_atomic
    from quantum_simulations_b200.storage.manifest import Manifest, read_manifest, write_manifest_atomic
    from quantum_simulations_b200.storage.pinned import PinnedBuffer
    from quantum_simulations_b200.wal.wal import WAL

    cd = validate_circuit_dict(circuit_dict)
    n = cd["number_of_qubits"]
    sim = ShardedSimulator(n, dtype)
    try:
        n_loc = n - sim.g
        chunk_size = min(chunk_size, 1 << n_loc)
        if (1 << n_loc) % chunk_size:
            raise ValueError("2^n must be divisible by chunk_size")
        work = Path(work_dir)
        levels = [st_["local_ops"] + st_["nonlocal_ops"] for st_ in build_steps(cd, n, False)]   # one entry per level
        wal = WAL(work / "wal.json", circuit_dict=cd) if (use_wal and sim.rank == 0) else None
        current, start = (wal.committed_buf, wal.done_steps) if wal else ("a", 0)
        if sim.dist is not None:
            current, start = sim.dist.broadcast_object((current, start), src=0)
        per_rank = (1 << n_loc) // chunk_size
        st, np_dtype = sim.shard.state, sim.dtype
        if start >= len(levels) and (start > 0 or not levels):
            if (_buf_dir(work, current) / "manifest.json").exists():
                return _buf_dir(work, current)                  # the committed buffer already holds the final state
            start = 0
        if start > 0:                                           # resume: committed chunks -> shards (identity layout)
            src = _buf_dir(work, current)
            m = read_manifest(src)
            if m.n_qubits != n or m.n_chunks * m.chunk_size != (1 << n):
                raise ValueError("the committed checkpoint does not belong to this circuit size")
            per_src = (1 << n_loc) // m.chunk_size
            for c in range(per_src):
                data = read_chunk(src / "chunks" / m.chunks[sim.rank * per_src + c], np.dtype(m.dtype))
                st.upload(data.astype(np_dtype, copy=False), c * m.chunk_size)
        every = checkpoint_every if checkpoint_every > 0 else max(len(levels) - start, 1)
        log_to_phys = None
        if use_staging:
            if staging_method not in ("heuristic", "greedy", "ilp"):
                raise ValueError(f"unknown staging method: {staging_method!r}")
            start, every = 0, max(len(levels), 1)               # one segment, planned by atlas_stages
        n_ckpt = 0
        lo = start
        while True:
            hi = min(lo + every, len(levels))
            final = hi >= len(levels)
            ops = [op for lv in levels[lo:hi] for op in lv]
            from_zero = lo == 0
            # only the LAST segment may leave an X pending on a rank bit as a renaming of the shards; a segment
            # that is continued must end with every amplitude where the identity layout puts it
            if use_staging:
                kw_ = dict(compiler_kw)
                kw_.setdefault("swap_anywhere", bool(sim.peer_swap))
                prog, log_to_phys = sharding.plan_atlas(cd, n_loc, sim.dtype.name, staging_method, True, **kw_)
            else:
                prog = sim.plan_ops(ops, zero_init=from_zero, **dict(compiler_kw, rank_flips=bool(final and compiler_kw.get("rank_flips", True))))
            if from_zero or ops:
                sim.run(prog, init=from_zero)
            logical = sim.rank ^ prog.rank_flip_mask
            dst = _buf_dir(work, _other(current))
            if sim.rank == 0:
                _wipe_buf(dst)
            if sim.dist is not None:
                sim.dist.barrier()
            first = logical * per_rank                          # logical shard order = index order
            bufs = [PinnedBuffer(chunk_size * np_dtype.itemsize), PinnedBuffer(chunk_size * np_dtype.itemsize)]
            try:
                st._ck(st.lib.qsv_download_async(st._h, bufs[0].ptr, 0, chunk_size))
                for c in range(per_rank):
                    st.sync()
                    if c + 1 < per_rank:
                        st._ck(st.lib.qsv_download_async(st._h, bufs[(c + 1) & 1].ptr, (c + 1) * chunk_size, chunk_size))
                    write_chunk_atomic(dst / "chunks" / chunk_filename(first + c), bufs[c & 1].array(np_dtype, chunk_size), np_dtype)
                st.sync()
            finally:
                for b in bufs:
                    b.free()
            if sim.dist is not None:
                sim.dist.barrier()                              # every chunk of every rank is durable
            if sim.rank == 0:
                total = per_rank * sim.world
                write_manifest_atomic(dst, Manifest(n_qubits=n, chunk_size=chunk_size, n_chunks=total, dtype=np_dtype.name,
                                                    chunks=[chunk_filename(i) for i in range(total)]))
                if log_to_phys is not None:
                    import json as _json
                    from quantum_simulations_b200.storage._atomic import publish_text
                    publish_text(work / "qubit_mapping.json", _json.dumps(list(log_to_phys)))
                if wal:
                    wal.commit_step(max(hi - 1, 0), _other(current))
            current = _other(current)
            n_ckpt += 1
            if sim.dist is not None:
                sim.dist.barrier()
            if final or (stop_after_checkpoints is not None and n_ckpt >= stop_after_checkpoints):
                break
            lo = hi
        if wal:
            wal.close()
        return _buf_dir(work, current)
    finally:
        sim.close()
