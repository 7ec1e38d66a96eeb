"""Host plumbing of the multi-GPU runner: rendezvous and small collectives over plain TCP (stdlib only).

One process per GPU is launched by ``python -m torch.distributed.run`` (or by anything else that sets RANK,
WORLD_SIZE, MASTER_ADDR, MASTER_PORT); the product code needs from the host side only
    * the 128-byte NCCL unique id of rank 0 and the 64-byte CUDA IPC handles of all ranks (bytes),
    * a collective AND ("do all ranks take the same path?"), barriers, the max of a timing,
    * sums of small float64 vectors (marginals, the sampler's leaf sums).
None of that is worth a tensor library: rank 0 listens on a port derived from MASTER_PORT, every other
rank keeps one connection to it, and every collective is "gather the encoded values on rank 0, send the
list back".  No state amplitude ever passes through here — the data path is NVLink (csrc/xchg.cuh).

Wire format: values are encoded by the closed codec below (None, bool, int, float, str, bytes, list, tuple, dict,
numeric ndarray) — NOT pickle: decoding a message can only ever build those types, so a process that reaches the
port cannot make a rank execute anything.
"""
from __future__ import annotations

import os
import socket
import struct
import time

import numpy as np

_MAGIC = b"QSVP1"
PORT_OFFSETS = (211, 422, 633, 844, 1055)        # candidates above MASTER_PORT (torchrun's own store sits ON it)


_MAX_MESSAGE = 1 << 34            # a length prefix beyond this is a protocol error, not an allocation request
_MAX_DEPTH = 32


def encode(obj, out: bytearray | None = None, depth: int = 0) -> bytearray:
    """Closed, typed encoding of the values the collectives carry (see module docstring)."""
    out = bytearray() if out is None else out
    if depth > _MAX_DEPTH:
        raise ValueError("host plumbing: value nested too deeply")
    if obj is None:
        out += b"N"
    elif isinstance(obj, (bool, np.bool_)):
        out += b"T" if obj else b"F"
    elif isinstance(obj, (int, np.integer)):
        raw = int(obj).to_bytes((int(obj).bit_length() + 8) // 8, "little", signed=True)
        out += b"I" + struct.pack("<I", len(raw)) + raw
    elif isinstance(obj, (float, np.floating)):
        out += b"D" + struct.pack("<d", float(obj))
    elif isinstance(obj, str):
        raw = obj.encode("utf-8")
        out += b"S" + struct.pack("<Q", len(raw)) + raw
    elif isinstance(obj, (bytes, bytearray, memoryview)):
        raw = bytes(obj)
        out += b"B" + struct.pack("<Q", len(raw)) + raw
    elif isinstance(obj, np.ndarray):
        if obj.dtype.kind not in "biufc":
            raise TypeError(f"host plumbing: arrays of dtype {obj.dtype} are not carried")
        dt = obj.dtype.str.encode("ascii")
        out += b"A" + struct.pack("<BB", len(dt), obj.ndim) + dt + struct.pack(f"<{obj.ndim}Q", *obj.shape)
        raw = np.ascontiguousarray(obj).tobytes()
        out += struct.pack("<Q", len(raw)) + raw
    elif isinstance(obj, (list, tuple)):
        out += (b"L" if isinstance(obj, list) else b"U") + struct.pack("<Q", len(obj))
        for x in obj:
            encode(x, out, depth + 1)
    elif isinstance(obj, dict):
        out += b"M" + struct.pack("<Q", len(obj))
        for k, v in obj.items():
            if not isinstance(k, (str, int, bool, float, bytes, tuple, np.integer)):
                raise TypeError(f"host plumbing: dict key of type {type(k).__name__} is not carried")
            encode(k, out, depth + 1)
            encode(v, out, depth + 1)
    elif isinstance(obj, np.complexfloating):
        encode(np.asarray(obj), out, depth)
    else:
        raise TypeError(f"host plumbing: values of type {type(obj).__name__} are not carried")
    return out


def decode(buf: bytes):
    """Inverse of encode; raises ValueError on anything malformed (truncated, unknown tag, trailing bytes)."""
    view = memoryview(buf)

    def take(at: int, n: int) -> tuple:
        if n < 0 or at + n > len(view):
            raise ValueError("host plumbing: truncated message")
        return view[at: at + n], at + n

    def one(at: int, depth: int):
        if depth > _MAX_DEPTH:
            raise ValueError("host plumbing: value nested too deeply")
        tag, at = take(at, 1)
        tag = bytes(tag)
        if tag == b"N":
            return None, at
        if tag in (b"T", b"F"):
            return tag == b"T", at
        if tag == b"I":
            n_, at = take(at, 4)
            raw, at = take(at, struct.unpack("<I", n_)[0])
            return int.from_bytes(raw, "little", signed=True), at
        if tag == b"D":
            raw, at = take(at, 8)
            return struct.unpack("<d", raw)[0], at
        if tag in (b"S", b"B"):
            n_, at = take(at, 8)
            raw, at = take(at, struct.unpack("<Q", n_)[0])
            return (bytes(raw).decode("utf-8") if tag == b"S" else bytes(raw)), at
        if tag == b"A":
            hd, at = take(at, 2)
            n_dt, ndim = struct.unpack("<BB", hd)
            dt_raw, at = take(at, n_dt)
            try:
                dt = np.dtype(bytes(dt_raw).decode("ascii"))
            except (TypeError, UnicodeDecodeError) as e:
                raise ValueError("host plumbing: bad array dtype") from e
            if dt.kind not in "biufc" or ndim > 16:
                raise ValueError("host plumbing: bad array header")
            sh, at = take(at, 8 * ndim)
            shape = struct.unpack(f"<{ndim}Q", sh)
            n_, at = take(at, 8)
            n_raw = struct.unpack("<Q", n_)[0]
            count = 1
            for d in shape:
                count *= d
            if count * dt.itemsize != n_raw:
                raise ValueError("host plumbing: array size does not match its shape")
            raw, at = take(at, n_raw)
            return np.frombuffer(bytes(raw), dtype=dt).reshape(shape).copy(), at
        if tag in (b"L", b"U"):
            n_, at = take(at, 8)
            n_items = struct.unpack("<Q", n_)[0]
            if n_items > len(view) - at:
                raise ValueError("host plumbing: truncated message")
            items = []
            for _ in range(n_items):
                x, at = one(at, depth + 1)
                items.append(x)
            return (items if tag == b"L" else tuple(items)), at
        if tag == b"M":
            n_, at = take(at, 8)
            n_items = struct.unpack("<Q", n_)[0]
            if n_items > len(view) - at:
                raise ValueError("host plumbing: truncated message")
            d = {}
            for _ in range(n_items):
                k, at = one(at, depth + 1)
                v, at = one(at, depth + 1)
                d[k] = v
            return d, at
        raise ValueError(f"host plumbing: unknown tag {tag!r}")

    obj, end = one(0, 0)
    if end != len(view):
        raise ValueError("host plumbing: trailing bytes after the value")
    return obj


def _send(sock: socket.socket, obj) -> None:
    data = encode(obj)
    sock.sendall(struct.pack("<Q", len(data)) + data)


def _recv_exact(sock: socket.socket, n: int) -> bytes:
    buf = bytearray()
    while len(buf) < n:
        part = sock.recv(min(n - len(buf), 1 << 20))
        if not part:
            raise ConnectionError("host plumbing: peer closed the connection")
        buf += part
    return bytes(buf)


def _recv(sock: socket.socket):
    (n,) = struct.unpack("<Q", _recv_exact(sock, 8))
    if n > _MAX_MESSAGE:
        raise ConnectionError(f"host plumbing: message of {n} bytes announced — not this protocol")
    return decode(_recv_exact(sock, n))


class HostPlumbing:
    """Star of TCP connections around rank 0.  Every method is a COLLECTIVE: all ranks call it, in the same order."""

    def __init__(self, rank: int, world: int, addr: str = "127.0.0.1", base_port: int = 29531, timeout: float = 300.0):
        self.rank, self.world = rank, world
        self._peers: list = []           # rank 0: sockets by rank (index 0 unused)
        self._up = None                  # other ranks: the socket to rank 0
        if world == 1:
            return
        ports = [base_port + o for o in PORT_OFFSETS]
        if rank == 0:
            srv = None
            for p in ports:
                try:
                    srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
                    srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
                    srv.bind((addr, p))
                    break
                except OSError:
                    srv.close()
                    srv = None
            if srv is None:
                raise RuntimeError(f"host plumbing: none of the ports {ports} on {addr} is free")
            srv.listen(world)
            srv.settimeout(timeout)
            self._peers = [None] * world
            while any(s is None for s in self._peers[1:]):
                conn, _ = srv.accept()
                conn.settimeout(timeout)
                try:
                    hello = _recv_exact(conn, len(_MAGIC) + 8)
                except (ConnectionError, socket.timeout):
                    conn.close()
                    continue
                r, w = struct.unpack("<II", hello[len(_MAGIC):])
                if hello[: len(_MAGIC)] != _MAGIC or w != world or not 0 < r < world or self._peers[r] is not None:
                    conn.close()
                    continue
                conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                conn.sendall(_MAGIC)
                self._peers[r] = conn
            srv.close()
        else:
            deadline = time.monotonic() + timeout
            while self._up is None:
                for p in ports:
                    try:
                        s = socket.create_connection((addr, p), timeout=5.0)
                        s.settimeout(timeout)
                        s.sendall(_MAGIC + struct.pack("<II", rank, world))
                        if _recv_exact(s, len(_MAGIC)) == _MAGIC:
                            s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                            self._up = s
                            break
                        s.close()
                    except (OSError, ConnectionError):
                        pass
                if self._up is None:
                    if time.monotonic() > deadline:
                        raise TimeoutError(f"host plumbing: rank {rank} could not reach rank 0 at {addr}:{ports}")
                    time.sleep(0.05)

    # ---- collectives ----
    def all_gather_object(self, obj) -> list:
        """[value of rank 0, ..., value of rank world-1] on every rank."""
        if self.world == 1:
            return [obj]
        if self.rank == 0:
            box = [obj] + [None] * (self.world - 1)
            for r in range(1, self.world):
                box[r] = _recv(self._peers[r])
            for r in range(1, self.world):
                _send(self._peers[r], box)
            return box
        _send(self._up, obj)
        return _recv(self._up)

    def broadcast_object(self, obj, src: int = 0):
        return self.all_gather_object(obj if self.rank == src else None)[src]

    def barrier(self) -> None:
        self.all_gather_object(None)

    def all(self, flag: bool) -> bool:
        return all(self.all_gather_object(bool(flag)))

    def allreduce(self, value, op: str = "sum"):
        """Elementwise sum / max / min of a float (returns float) or of an array (returns float64 ndarray).
        The reduction runs in rank order on every rank: bit-identical results everywhere."""
        scalar = np.isscalar(value)
        parts = self.all_gather_object(np.asarray(value, dtype=np.float64))
        acc = np.array(parts[0], dtype=np.float64, copy=True)
        for p in parts[1:]:
            if op == "sum":
                acc = acc + p
            elif op == "max":
                acc = np.maximum(acc, p)
            elif op == "min":
                acc = np.minimum(acc, p)
            else:
                raise ValueError(f"allreduce: unknown op {op!r}")
        return float(acc) if scalar else acc

    def close(self) -> None:
        for s in self._peers:
            if s is not None:
                s.close()
        if self._up is not None:
            self._up.close()
        self._peers, self._up = [], None


_plumbing: HostPlumbing | None = None


def init_plumbing() -> HostPlumbing:
    """The process-wide plumbing object, from RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun sets them)."""
    global _plumbing
    if _plumbing is None:
        rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
        port = int(os.environ.get("QSV_PLUMBING_PORT", os.environ.get("MASTER_PORT", "29531")))
        _plumbing = HostPlumbing(rank, world, addr, port)
    return _plumbing


def shutdown_plumbing() -> None:
    global _plumbing
    if _plumbing is not None:
        _plumbing.close()
        _plumbing = None
