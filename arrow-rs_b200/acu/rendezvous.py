"""One-process-per-GPU plumbing for the benches: exchange the NCCL unique id, barrier, and
max-over-ranks of a host float, for a torchrun-style launch (RANK / LOCAL_RANK / WORLD_SIZE /
MASTER_ADDR / MASTER_PORT in the environment).

Two transports for the 128-byte id (ACU_RENDEZVOUS=socket|torch, default socket):
  * socket — a plain TCP exchange on MASTER_PORT+17, no torch import at all (the data path never
             needs torch: every collective of the hot path is acu_comm_allreduce_* over the
             library's own NCCL communicator, and a rank starts ~10 s sooner without the import);
  * torch  — torch.distributed (backend nccl) broadcast. torch must be imported BEFORE the library
             opens NCCL so that both use the copy bundled with torch.
Barrier and max-over-ranks always go through the library's communicator."""
import ctypes as C
import os
import socket
import struct
import time

from . import _abi as abi


class Group:
    def __init__(self, ctx, rank, local_rank, world):
        self.ctx, self.rank, self.local_rank, self.world = ctx, rank, local_rank, world
        self._torch_dist = None
        if world <= 1:
            return
        lib, h = ctx.lib, ctx.h
        mode = os.environ.get("ACU_RENDEZVOUS", "socket")
        if mode == "torch":
            import torch  # noqa: F401  (first: its bundled NCCL must be the one already loaded when the library dlopens it)
            import torch.distributed as dist
        idb = (C.c_uint8 * abi.NCCL_UNIQUE_ID_BYTES)()
        if rank == 0:
            assert lib.acu_comm_get_unique_id(idb) == abi.OK
        if mode != "torch":
            raw = self._socket_broadcast(bytes(idb))
            idb = (C.c_uint8 * abi.NCCL_UNIQUE_ID_BYTES)(*raw)
        else:
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            t = torch.tensor(list(idb), dtype=torch.uint8, device="cuda")
            dist.broadcast(t, 0)
            idb = (C.c_uint8 * abi.NCCL_UNIQUE_ID_BYTES)(*t.cpu().tolist())
            self._torch_dist = dist
        ctx.check(lib.acu_comm_init(h, idb, rank, world))

    @staticmethod
    def _recv_exact(sock, n):
        """Read exactly n bytes or raise: recv() returning b'' means the peer closed the connection."""
        buf = b""
        while len(buf) < n:
            chunk = sock.recv(n - len(buf))
            if not chunk:
                raise ConnectionError(f"rendezvous peer closed the connection after {len(buf)} of {n} bytes")
            buf += chunk
        return buf

    def _socket_broadcast(self, payload, timeout=120.0):
        """Rank 0 hands the NCCL unique id to every other rank over TCP. A client introduces itself with
        (job nonce, rank); rank 0 answers each DISTINCT valid rank once and ignores stray connections, so a port scanner or a
        stale process cannot consume a slot. Every wait is bounded by `timeout` seconds."""
        addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
        port = int(os.environ.get("MASTER_PORT", "29500")) + 17
        nonce = (os.environ.get("TORCHELASTIC_RUN_ID", "") + ":" + os.environ.get("MASTER_PORT", "29500")).encode()[:64].ljust(64, b"\0")
        if self.rank == 0:
            srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            srv.bind((addr, port))
            srv.listen(self.world + 8)
            deadline = time.time() + timeout
            served = set()
            try:
                while len(served) < self.world - 1:
                    left = deadline - time.time()
                    if left <= 0:
                        raise TimeoutError(f"rendezvous: only ranks {sorted(served)} of {self.world - 1} peers connected within {timeout:.0f} s")
                    srv.settimeout(left)
                    try:
                        conn, _ = srv.accept()
                    except socket.timeout:
                        continue
                    try:
                        conn.settimeout(5.0)
                        hello = self._recv_exact(conn, 68)
                        peer = struct.unpack("<I", hello[64:])[0]
                        if hello[:64] == nonce and 0 < peer < self.world and peer not in served:
                            conn.sendall(struct.pack("<I", len(payload)) + payload)
                            served.add(peer)
                    except (OSError, ConnectionError):
                        pass  # a stray or broken connection does not consume a slot
                    finally:
                        conn.close()
            finally:
                srv.close()
            return payload
        deadline = time.time() + timeout
        while True:
            try:
                s = socket.create_connection((addr, port), timeout=5)
                s.settimeout(max(deadline - time.time(), 1.0))
                s.sendall(nonce + struct.pack("<I", self.rank))
                n = struct.unpack("<I", self._recv_exact(s, 4))[0]
                data = self._recv_exact(s, n)
                s.close()
                return data
            except (OSError, ConnectionError):
                if time.time() > deadline:
                    raise
                time.sleep(0.1)

    def barrier(self):
        self.ctx.sync()
        if self.world > 1:
            v = (C.c_int64 * 1)(1)
            self.ctx.check(self.ctx.lib.acu_comm_allreduce_i64_sum(self.ctx.h, v, 1))
            assert v[0] == self.world

    def max_over_ranks(self, x):
        """max of a non-negative host float over all ranks (device times: the slowest rank counts)."""
        if self.world <= 1:
            return float(x)
        import numpy as np
        bits = (C.c_uint64 * 1)(int(np.array([x], dtype=np.float64).view(np.uint64)[0]))
        cnt = (C.c_int64 * 1)(1)
        self.ctx.check(self.ctx.lib.acu_comm_allreduce_aggregates(self.ctx.h, abi.F64, abi.MAX, bits, cnt, 1))
        return float(np.array([bits[0]], dtype=np.uint64).view(np.float64)[0])

    def close(self):
        if self.world > 1:
            self.ctx.lib.acu_comm_destroy(self.ctx.h)
        if self._torch_dist is not None:
            self._torch_dist.destroy_process_group()
            self._torch_dist = None
