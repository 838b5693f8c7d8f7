"""Multi-GPU plumbing for independent stereo sequences ("replicas only").

A VO frame does not shard and the pnpmatch stage is sequential in time inside one sequence
(src/Tracking.cc:107-121), so the only parallel axis is the sequence: every rank owns a disjoint
set of sequences and runs the whole front-end on them.  There is no data-path collective.
torch.distributed (NCCL on the GPU box, gloo in the CPU tests) is used for exactly three things:
the start/stop barrier of a timed region, the max-over-ranks of the device time and the sum of
the frames every rank processed.
"""
import os


def env_rank_world():
    """(rank, world, local_rank) as torchrun exports them; (0, 1, 0) outside torchrun."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def assign_sequences(n_sequences, world, rank):
    """Sequence ids owned by `rank`: a contiguous block, sizes differing by at most one, every id
    owned by exactly one rank (BASELINE.json configs[4]: 8 sequences over 1/2/4/8 GPUs)."""
    if world < 1 or not (0 <= rank < world) or n_sequences < 0:
        raise ValueError("bad shard request: %d sequences, rank %d of %d" % (n_sequences, rank, world))
    base, extra = divmod(n_sequences, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def bind_near_gpu(index, pci_bus_id=None):
    """Pin this process to the CPU cores NVML reports as local to the GPU (its NUMA node / PCIe root), so the pinned
    host buffers allocated afterwards are placed there and the per-step H2D/D2H copies of eight replicas do not all
    cross the socket interconnect.  Placement only: nothing on the data path changes.  Returns the core list, or
    None when NVML or the affinity call is unavailable (single-node VMs report every core: a no-op)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByPciBusId(pci_bus_id) if pci_bus_id else pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * i + b for i, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = sorted(c for c in cpus if c in allowed)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


class Group:
    """Thin wrapper over a torch.distributed process group (or nothing when world == 1)."""

    def __init__(self, backend=None, device=None):
        self.rank, self.world, self.local_rank = env_rank_world()
        self.device = device
        self.dist = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            if not dist.is_initialized():
                kw = {}
                if backend == "nccl" and device is not None:
                    kw["device_id"] = torch.device("cuda", device)
                # NCCL prints its version banner on the process's stdout when the communicator is created; bench.py's
                # stdout must carry exactly one JSON line, so fd 1 points at stderr while the group comes up
                import sys
                sys.stdout.flush()
                saved = os.dup(1)
                os.dup2(2, 1)
                try:
                    dist.init_process_group(backend or "gloo", **kw)
                    if (backend or "gloo") == "nccl":
                        torch.cuda.synchronize()
                        dist.barrier()          # forces communicator creation now
                        torch.cuda.synchronize()
                finally:
                    sys.stdout.flush()
                    os.dup2(saved, 1)
                    os.close(saved)
            self.dist = dist
            self.backend = dist.get_backend()
        else:
            self.backend = None

    def _tensor(self, values):
        import torch
        dev = "cuda" if self.backend == "nccl" else "cpu"
        return torch.tensor(values, dtype=torch.float64, device=dev)

    def barrier(self):
        """Device-synchronising barrier that brackets a timed region."""
        if self.backend == "nccl" or (self.dist is None and self.device is not None):
            import torch
            torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            if self.backend == "nccl":
                import torch
                torch.cuda.synchronize()

    def max_over_ranks(self, values):
        """Element-wise maximum of a list of per-rank times (ms)."""
        if self.dist is None:
            return [float(v) for v in values]
        t = self._tensor(values)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def sum_over_ranks(self, values):
        if self.dist is None:
            return [float(v) for v in values]
        t = self._tensor(values)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.tolist()

    def aggregate_fps(self, frames_this_rank, ms_this_rank):
        """Whole-job throughput: frames of all ranks / slowest rank's device time."""
        frames = self.sum_over_ranks([frames_this_rank])[0]
        ms = self.max_over_ranks([ms_this_rank])[0]
        return frames / (ms * 1e-3), frames, ms

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()
            self.dist = None
