"""Host logic of the sharded frame-to-map Hamming search (north_star multi-GPU piece).

The accumulated map descriptors (what Map::getKeypoints, src/mymap.cpp:28-74, hands to
featureMatching, src/lidar_odometry.cpp:197-206) are split into contiguous index ranges, one per
rank.  Per call: every rank matches the (replicated) query set against its resident shard with
GLOBAL target indices, the per-rank candidate records {k1, k2, rq} (24 B/query) are all-gathered,
and a merge by (distance, global index) reproduces the single-GPU first-minimum rule bit for bit.
`rq` (best query for the rank's best target) rides along, so the mutual check needs no second
exchange.

The compute back end is injected: production uses the C ABI on the GPU (gpu_backend), the CPU
gloo tests inject an oracle-based back end from tests/ -- this module never imports the oracle.
"""
import numpy as np

CAND_DTYPE = np.dtype([("k1", "<u8"), ("k2", "<u8"), ("rq", "<u4"), ("pad", "<u4")])
NONE_KEY = np.uint64(0xFFFFFFFFFFFFFFFF)


def shard_range(total, world, rank):
    """contiguous split: rank r holds [r * ceil(T/G), min(T, (r+1) * ceil(T/G)))"""
    per = (total + world - 1) // world
    lo = min(total, rank * per)
    return lo, min(total, lo + per)


def merge_records(gathered):
    """numpy mirror of merge_cands_kernel (csrc/hamming.cu): gathered (ranks, Q) records -> (Q,)"""
    g = np.asarray(gathered)
    ranks, q = g.shape
    keys = np.concatenate([g["k1"], g["k2"]], axis=0)                  # (2R, Q)
    order = np.sort(keys, axis=0)
    out = np.zeros(q, CAND_DTYPE)
    out["k1"], out["k2"] = order[0], order[1] if 2 * ranks > 1 else NONE_KEY
    win = np.argmin(g["k1"], axis=0)                                   # first minimum = lowest rank
    out["rq"] = g["rq"][win, np.arange(q)]
    out["rq"][out["k1"] == NONE_KEY] = 0xFFFFFFFF
    return out


class ShardedMap:
    """one rank's view of the sharded map"""

    def __init__(self, backend, rank=0, world=1, all_gather=None):
        self.backend, self.rank, self.world = backend, rank, world
        self.all_gather = all_gather            # callable(local (Q,) records) -> (world, Q) records
        self.lo = self.hi = 0

    def load(self, global_map):
        """keep this rank's contiguous shard of the global descriptor array resident"""
        self.lo, self.hi = shard_range(len(global_map), self.world, self.rank)
        self.backend.set_shard(global_map[self.lo:self.hi])

    def match(self, queries):
        local = self.backend.match_shard(queries, self.lo)             # (Q,) records, global indices
        if self.world == 1 or self.all_gather is None:
            return self.backend.merge(local[None, :])
        return self.backend.merge(self.all_gather(local))

    def match_sharded_reverse(self, queries, all_reduce_max):
        """Variant that keeps the reverse (best query per winning target) pass sharded: per-rank records
        carry no rq; after the merge every rank searches only the winners that live in ITS shard
        (Q*Q/ranks pairs) and the int32 results (-1 = not mine) are combined by an all-reduce(MAX) --
        exactly one rank owns each winner."""
        local = self.backend.match_shard(queries, self.lo, with_rq=False)
        merged = self.backend.merge(self.all_gather(local) if self.world > 1 else local[None, :])
        rq = self.backend.reverse_owned(queries, self.lo, merged)      # (Q,) int32, -1 where not owned
        if self.world > 1:
            rq = all_reduce_max(rq)
        merged = merged.copy()
        merged["rq"] = np.where(merged["k1"] != NONE_KEY, rq.astype(np.int64) & 0xFFFFFFFF, 0xFFFFFFFF).astype(np.uint32)
        return merged

    @staticmethod
    def correspondences(merged):
        """mutual-NN filter (src/lidar_odometry.cpp:234-242) from merged records"""
        q = np.arange(len(merged))
        keep = (merged["k1"] != NONE_KEY) & (merged["rq"] == q)
        idx = (merged["k1"] & np.uint64(0xFFFFFFFF)).astype(np.int64)
        return np.stack([q[keep], idx[keep]], axis=1).astype(np.int32)


class GpuBackend:
    """production back end: bshot_map_append / bshot_match_map / bshot_merge_cands_dev (C ABI)"""

    def __init__(self, ctx):
        self.ctx = ctx

    def set_shard(self, desc):
        self.ctx.map_reset()
        self.ctx.map_append(desc)

    def match_shard(self, queries, global_base, with_rq=True):
        if with_rq:
            return self.ctx.match_map(queries, global_base)
        import torch
        q = np.ascontiguousarray(queries, dtype=np.uint64).reshape(-1, 6)
        dq = torch.from_numpy(q.view(np.int64)).cuda()
        out = torch.empty((len(q), 3), dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        self.ctx.match_shard_dev(dq.data_ptr(), len(q), global_base, False, out.data_ptr())
        self.ctx.sync()
        return out.cpu().numpy().view(CAND_DTYPE).reshape(len(q))

    def reverse_owned(self, queries, global_base, merged):
        import torch
        q = np.ascontiguousarray(queries, dtype=np.uint64).reshape(-1, 6)
        dq = torch.from_numpy(q.view(np.int64)).cuda()
        dm = torch.from_numpy(np.ascontiguousarray(merged).view(np.int64).reshape(len(q), 3)).cuda()
        rq = torch.empty(len(q), dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        self.ctx.reverse_owned_dev(dq.data_ptr(), len(q), global_base, dm.data_ptr(), rq.data_ptr())
        self.ctx.sync()
        return rq.cpu().numpy()

    def merge(self, gathered):
        import torch
        g = np.ascontiguousarray(gathered)
        ranks, q = g.shape
        d = torch.from_numpy(g.view(np.int64).reshape(ranks, q, 3)).cuda()
        out = torch.empty((q, 3), dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        self.ctx.merge_cands_dev(d.data_ptr(), ranks, q, out.data_ptr())
        self.ctx.sync()
        return out.cpu().numpy().view(CAND_DTYPE).reshape(q)


class DeviceShardedMatcher:
    """Device-resident sharded frame-to-map match of one rank: queries, records and the exchange stay on the GPU and
    everything is stream-ordered on the context stream (no host synchronisation per call).

    Per call at world > 1: shard search (no rq) -> exchange of the 24 B/query records -> merge by (distance, global
    index) -> reverse pass only for the winners this rank owns (Q * Q / ranks pairs) -> exchange of the 4 B/query
    reverse result -> records completed.  Three interchangeable exchanges:
      * "ipc" (default): the communicator of the C ABI (bshot_comm_*, include/bshot_b200.h): CUDA-IPC peer memory, the
        whole call is bshot_match_map_sharded_dev -- six kernel launches with the stores and the flag barriers fused into
        the merge kernels.  torch.distributed only carries the 64-byte handles once at set-up;
      * "peer": gather buffer, rq array and a flag array are torch symmetric-memory allocations mapped on every rank
        over NVLink / NVSwitch; the records and the owners' reverse results are STORED straight into every rank's
        buffers (bshot_push_cands_dev / bshot_reverse_owned_push_dev) and bshot_peer_barrier_dev replaces each
        collective -- no NCCL call on the data path;
      * "nccl": one all_gather_into_tensor of the records and one all_reduce(MAX) of the reverse result.
    mode="auto" takes "peer" when symmetric memory can be set up on every rank, else "nccl"."""

    def __init__(self, ctx, world, rank, nq, mode="auto", device=None):
        import torch
        self.torch = torch
        self.ctx, self.world, self.rank, self.nq = ctx, world, rank, nq
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.stream = torch.cuda.ExternalStream(ctx.stream)
        self.cand = torch.empty((nq, 3), dtype=torch.int64, device=dev)
        self.merged = torch.empty((nq, 3), dtype=torch.int64, device=dev)
        self.peer = None
        self.exchange = "none"
        self.ipc = False
        if world == 1:
            return
        import torch.distributed as dist
        self.dist = dist
        if mode in ("auto", "ipc"):
            ok = 1
            try:
                ctx.comm_create(rank, world, nq)
                mine = torch.from_numpy(ctx.comm_export()).to(dev)
                allh = torch.empty((world, 64), dtype=torch.uint8, device=dev)
                dist.all_gather_into_tensor(allh.view(-1), mine)
                ctx.comm_import(allh.cpu().numpy())
            except Exception as e:  # noqa: BLE001 -- e.g. no peer access between the GPUs
                ok, self.ipc_error = 0, e
            t = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)           # every rank takes the same path
            torch.cuda.synchronize()
            if int(t.item()) == 1:
                self.ipc = True
                self.exchange = "ipc"
                return
            if mode == "ipc":
                raise RuntimeError(f"CUDA IPC communicator unavailable on some rank ({getattr(self, 'ipc_error', None)})")
        if mode in ("auto", "peer"):
            err = None
            try:
                import torch.distributed._symmetric_memory as symm_mem
                g = symm_mem.empty(world * nq * 3, dtype=torch.int64, device=dev)
                r = symm_mem.empty(nq, dtype=torch.int32, device=dev)
                f = symm_mem.empty(64, dtype=torch.int32, device=dev)
                g.zero_(); r.fill_(-1); f.zero_()
                torch.cuda.synchronize()
                ctx.peer_barrier_reset()   # the barrier epoch belongs to this (zeroed) flag array
                hs = [symm_mem.rendezvous(t, dist.group.WORLD) for t in (g, r, f)]
                self.peer = dict(g=g, r=r, f=f, handles=hs, pg=int(hs[0].buffer_ptrs_dev), pr=int(hs[1].buffer_ptrs_dev),
                                 pf=int(hs[2].buffer_ptrs_dev))
            except Exception as e:  # noqa: BLE001 -- any failure means "no symmetric memory here"
                err, self.peer = e, None
            ok = torch.tensor([1 if self.peer is not None else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)          # every rank takes the same path; also the barrier
            torch.cuda.synchronize()                           # that makes all flag arrays zero before first use
            if int(ok.item()) == 0:
                self.peer = None
                if mode == "peer":
                    raise RuntimeError(f"symmetric memory unavailable on some rank ({err})")
        if self.peer is not None:
            self.exchange = "peer"
        else:
            self.exchange = "nccl"
            self.gathered = torch.empty((world, nq, 3), dtype=torch.int64, device=dev)
            self.rq = torch.empty(nq, dtype=torch.int32, device=dev)

    def describe(self):
        return {"none": "none",
                "ipc": "per call: bshot_match_map_sharded_dev (C ABI): 6 launches; records and reverse results STORED into every rank's "
                       "CUDA-IPC peer buffers by the merge kernels, flag release / acquire fused into them (no NCCL on the data path)",
                "peer": "per call: peer-memory stores (24 B/query records to every rank, 4 B/query reverse result from the "
                        "owner) + 2 flag barriers over symmetric memory (no NCCL on the data path)",
                "nccl": "per call: nccl all_gather of 24 B/query records + all_reduce of 4 B/query reverse result"}[self.exchange]

    def match(self, d_q_ptr, global_base):
        """asynchronous; returns the device tensor (nq, 3) int64 = bshot_cand records with rq filled"""
        c, nq, w, rk = self.ctx, self.nq, self.world, self.rank
        if w == 1:
            c.match_shard_dev(d_q_ptr, nq, global_base, True, self.cand.data_ptr())
            c.merge_cands_dev(self.cand.data_ptr(), 1, nq, self.merged.data_ptr())
            return self.merged
        if self.ipc:
            c.match_map_sharded_dev(d_q_ptr, nq, global_base, self.merged.data_ptr())
            return self.merged
        c.match_shard_dev(d_q_ptr, nq, global_base, False, self.cand.data_ptr())
        if self.peer is not None:
            p = self.peer
            c.push_cands_dev(self.cand.data_ptr(), nq, p["pg"], w, rk)
            c.peer_barrier_dev(p["pf"], w, rk)
            c.merge_cands_dev(p["g"].data_ptr(), w, nq, self.merged.data_ptr())
            c.reverse_owned_push_dev(d_q_ptr, nq, global_base, self.merged.data_ptr(), p["pr"], w, rk)
            c.peer_barrier_dev(p["pf"], w, rk)
            c.apply_rq_dev(self.merged.data_ptr(), p["r"].data_ptr(), nq)
            return self.merged
        torch, dist = self.torch, self.dist
        with torch.cuda.stream(self.stream):
            dist.all_gather_into_tensor(self.gathered.view(-1), self.cand.view(-1))
        c.merge_cands_dev(self.gathered.data_ptr(), w, nq, self.merged.data_ptr())
        c.reverse_owned_dev(d_q_ptr, nq, global_base, self.merged.data_ptr(), self.rq.data_ptr())
        with torch.cuda.stream(self.stream):
            dist.all_reduce(self.rq, op=dist.ReduceOp.MAX)     # one owner per query, the others hold -1
        c.apply_rq_dev(self.merged.data_ptr(), self.rq.data_ptr(), nq)
        return self.merged

    def check(self):
        """raises if a peer barrier gave up waiting for a rank"""
        if self.ipc:
            self.ctx.comm_check()
        if self.peer is not None and self.ctx.peer_barrier_timeouts():
            raise RuntimeError("a peer barrier timed out (a rank did not arrive)")
