"""Data-parallel plumbing of the hot path (one process per GPU, torch.distributed).

The path shards two ways:
  * rollouts / Jacobians / reward-to-go are independent units -> contiguous block partition of the unit index,
    replicated actor weights, NO collective on the data path (``shard_range``);
  * the update is synchronous data parallelism: every rank takes B/world rows of the same globally sampled
    minibatch, computes gradient SUMS scaled by 1/B_global, and one all-reduce(sum) per network gives every
    replica the full-batch gradient before the identical Adam step (``allreduce_sum``).  Critic and actor
    steps are sequentially dependent (RL.py:104-109), hence one collective per network.
The functions take the process-group module/object so that the same code runs over NCCL on GPUs and over
gloo in the CPU tests.
"""


def shard_range(n, rank, world):
    """Contiguous block [lo, hi) of ``n`` units owned by ``rank``; the first n % world ranks get one extra."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum(tensor, dist):
    """In-place sum over the group (NCCL over NVLink on GPUs).  No-op without a group or with one rank."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return tensor
    dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


def global_batch(local_batch, dist):
    if dist is None or not dist.is_initialized():
        return int(local_batch)
    return int(local_batch) * dist.get_world_size()


# ------------------------------------------------------------------------------------------ NVLink peer memory
class PeerUnavailable(RuntimeError):
    """The GPUs / processes of this job cannot share memory through CUDA IPC (raised on every rank alike)."""


class PeerRegion:
    """One zero-filled cudaMalloc block per rank that every rank of the box has mapped (``cacto_peer_*`` of
    include/cacto_b200.h): ``bases[r]`` is the address of rank r's block in THIS process.  Built either across
    processes (``exchange``: CUDA-IPC handles travel through ``dist.all_gather_object``) or inside one process
    (``local_group``: the blocks of all simulated ranks on the current device -- what the single-GPU test uses)."""

    def __init__(self, nbytes, rank, world, bases, owned, opened):
        self.nbytes, self.rank, self.world, self.bases = int(nbytes), int(rank), int(world), list(bases)
        self._owned, self._opened = owned, opened

    @staticmethod
    def _alloc(nbytes):
        import ctypes as C
        from ._lib import check, lib
        p = C.c_void_p()
        check(lib.cacto_peer_alloc(int(nbytes), C.byref(p)), 'peer_alloc')
        return p.value

    @classmethod
    def exchange(cls, nbytes, dist):
        """Collective over ``dist``.  Every rank allocates and exports its block, the handles travel through
        ``all_gather_object``, every rank maps the others, and a final MIN all-reduce makes the outcome unanimous: if any
        rank could not export or map (IPC not permitted between the processes, no peer access between two GPUs), ALL ranks
        release what they hold and raise ``PeerUnavailable`` -- no rank is left waiting for a peer that gave up."""
        import ctypes as C
        import torch
        from ._lib import lib
        rank, world = dist.get_rank(), dist.get_world_size()
        base, handle, err = None, None, None
        try:
            base = cls._alloc(nbytes)
            buf = C.create_string_buffer(64)
            rc = lib.cacto_peer_export(C.c_void_p(base), buf)
            if rc != 0:
                raise RuntimeError('cacto_peer_export: CUDA error %d' % rc)
            handle = bytes(buf.raw)
        except RuntimeError as e:
            err = e
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        bases, opened = [], []
        if err is None and all(h is not None for h in handles):
            for r, h in enumerate(handles):
                if r == rank:
                    bases.append(base)
                    continue
                q = C.c_void_p()
                rc = lib.cacto_peer_open(C.create_string_buffer(h, 64), C.byref(q))
                if rc != 0:
                    err = RuntimeError('cacto_peer_open (rank %d): CUDA error %d' % (r, rc))
                    break
                bases.append(q.value)
                opened.append(q.value)
        elif err is None:
            err = RuntimeError('a peer could not export its block')
        ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=torch.device('cuda', torch.cuda.current_device()))
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        region = cls(nbytes, rank, world, bases, base, opened)
        if int(ok[0]) == 0:
            region.close()
            raise PeerUnavailable(str(err) if err is not None else 'another rank could not map the peer blocks')
        return region

    @classmethod
    def local_group(cls, nbytes, world):
        bases = [cls._alloc(nbytes) for _ in range(world)]
        return [cls(nbytes, r, world, bases, bases[r], []) for r in range(world)]

    def tensor(self, offset, count, dtype, device):
        """Zero-copy torch view of ``count`` elements at byte ``offset`` of this rank's own block."""
        import numpy as np
        import torch
        np_dt = np.dtype({torch.float32: 'f4', torch.int32: 'i4', torch.uint8: 'u1'}[dtype])
        assert offset + count * np_dt.itemsize <= self.nbytes

        class _View:
            pass
        v = _View()
        v.__cuda_array_interface__ = dict(shape=(int(count),), typestr=np_dt.str, data=(self.bases[self.rank] + int(offset), False),
                                          version=2, strides=None)
        v._keepalive = self
        t = torch.as_tensor(v, device=device)
        assert t.data_ptr() == self.bases[self.rank] + int(offset)
        return t

    def close(self):
        import ctypes as C
        from ._lib import lib
        for q in self._opened:
            lib.cacto_peer_close(C.c_void_p(q))
        self._opened = []
        if self._owned:
            lib.cacto_peer_free(C.c_void_p(self._owned))
            self._owned = None


class PeerReduce:
    """Gradient exchange of the data-parallel update without a collective launch: the gradient blocks of the critic and
    the actor and two rows of arrival words live in a ``PeerRegion``; ``table(net)`` gives the per-rank pointer arrays that
    ``cacto_adam_step_peer`` (update.cu: k_adam_peer) sums over NVLink inside the Adam kernel.  Layout of every rank's
    block (identical on all ranks): [critic flags 128 B][actor flags 128 B][critic grad][actor grad], 128-byte aligned.
    ``max_ctas`` bounds the CTAs of the kernel (0 = no bound); only ranks simulated on ONE device need it (the CTAs spin while
    they wait for the peers, and an SM full of spinning CTAs cannot be re-carved for the 200 KB gradient kernels)."""

    MAX_PEERS = 8

    @staticmethod
    def region_bytes(n_critic, n_actor):
        r = lambda b: (b + 127) // 128 * 128
        return 256 + r(4 * n_critic) + r(4 * n_actor)

    def __init__(self, region, critic, actor, max_ctas=0):
        import ctypes as C
        import torch
        if region.world > self.MAX_PEERS:
            raise ValueError('peer reduce supports up to %d GPUs of one box' % self.MAX_PEERS)
        assert region.nbytes >= self.region_bytes(critic.n, actor.n)
        self.region, self.rank, self.world = region, region.rank, region.world
        r = lambda b: (b + 127) // 128 * 128
        off = {id(critic): (0, 256), id(actor): (128, 256 + r(4 * critic.n))}
        self._tables = {}
        for net in (critic, actor):
            f_off, g_off = off[id(net)]
            grad = region.tensor(g_off, net.n, torch.float32, net.params.device)
            net.grad = grad                                  # the kernels accumulate straight into the shared block
            net._grad_views = net._make_views(grad)
            grads = (C.c_void_p * region.world)(*[b + g_off for b in region.bases])
            flags = (C.c_void_p * region.world)(*[b + f_off for b in region.bases])
            self._tables[id(net)] = (grads, flags, region.rank, int(max_ctas))

    def table(self, net):
        return self._tables[id(net)]


# ------------------------------------------------------------------------------------------ host placement
def bind_to_gpu_numa_node(device_index):
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off (``/sys/bus/pci/devices/<bdf>/numa_node``), so that
    the pinned host buffers of the warm-start path (RL_AC.rollout_to_host: 1 GB of fp64 trajectories per step and GPU) are
    first-touched on that node and the feeder threads run next to them.  Without it the 8 ranks of a box all allocate on node 0
    and share its memory controllers and one socket's PCIe root (round 1: 1.9 x e2e throughput from 1 to 8 GPUs).
    Returns a dict describing what was done (for the bench line); never raises -- placement is an optimisation."""
    import os
    info = {'device': int(device_index), 'numa_node': None, 'cpus': None, 'bound': False}
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bdf = '%04x:%02x:%02x.0' % (int(getattr(p, 'pci_domain_id', 0)), int(p.pci_bus_id), int(p.pci_device_id))
        with open(f'/sys/bus/pci/devices/{bdf}/numa_node') as f:
            node = int(f.read())
        info['numa_node'] = node
        if node < 0:
            return info
        cpus = set()
        with open(f'/sys/devices/system/node/node{node}/cpulist') as f:
            cpulist = f.read().strip()
        for part in cpulist.split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            info['cpus'] = '%d cpus of node %d' % (len(allowed), node)
            info['bound'] = True
            try:                                   # prefer the node for future allocations of this process (libnuma, if present)
                import ctypes
                numa = ctypes.CDLL('libnuma.so.1')
                if numa.numa_available() >= 0:
                    numa.numa_set_preferred(node)
                    info['preferred'] = True
            except OSError:
                pass
    except Exception as exc:                       # no sysfs entry (container), odd PCI ids: report and go on
        info['error'] = f'{type(exc).__name__}: {exc}'
    return info
