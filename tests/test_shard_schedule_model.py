"""CPU model check of the native row-sharded driver's SCHEDULE (csrc/shard.cu, fdr_shard_restore_rows): is every pair of
conflicting memory accesses -- on this rank or on a peer, within an image or across two consecutive images -- ordered by what
the driver actually issues (stream order, event record/wait, cross-rank barriers)?

The driver puts the passes of the three colour planes on a compute stream, the link kernels on a second and the barriers on a
third stream, with events in between; peers write into each other's slabs and staging planes.  A missing wait would be a
data race that the GPU tests see only once in a while.  Here the issue order of fdr_shard_restore_rows is restated op by op
(same order, same events: kinds 0..6 per unit + fork/join), every op is annotated with the regions it reads and writes
(layout of DESIGN.md section 2 / shard.cu: slab [unit][row block of rank q], staging [unit][column owner], Nyquist vectors,
raw planes, extrema), the happens-before relation is the transitive closure of the issued edges, and all conflicting pairs
must be ordered.  Negative controls drop one wait each and must be flagged.  Replaces nothing in the reference: its MPI
collectives are blocking (fft_mpi.cpp:170-279), which is exactly the serialisation this driver removes."""
import itertools

import pytest


class Schedule:
    def __init__(self, G, U, staged, drop=None):
        self.G, self.U, self.staged, self.drop = G, U, staged, drop
        self.ops = []          # (rank, stream, name, reads, writes)
        self.succ = {}         # op index -> set of successors
        self.last = {}         # (rank, stream) -> last op index
        self.event = {}        # (rank, event key) -> op index of the latest record
        self.barriers = {}     # (set, epoch) -> {rank: (arrive, leave)}
        self.epoch = {}

    def op(self, rank, stream, name, reads=(), writes=()):
        i = len(self.ops)
        self.ops.append((rank, stream, name, frozenset(reads), frozenset(writes)))
        self.succ[i] = set()
        prev = self.last.get((rank, stream))
        if prev is not None:
            self.succ[prev].add(i)
        self.last[(rank, stream)] = i
        return i

    def record(self, rank, key, stream):
        self.event[(rank, key)] = self.op(rank, stream, "record %s" % (key,))

    def wait(self, rank, key, stream):
        if self.drop == ("wait", key[0], stream):
            return
        w = self.op(rank, stream, "wait %s" % (key,))
        src = self.event.get((rank, key))
        if src is not None:                      # waiting on a never-recorded event is a no-op in CUDA
            self.succ[src].add(w)

    def barrier(self, rank, set_index, stream):
        ep = self.epoch.get((rank, set_index), 0) + 1
        self.epoch[(rank, set_index)] = ep
        arrive = self.op(rank, stream, "barrier %d.%d arrive" % (set_index, ep))
        leave = self.op(rank, stream, "barrier %d.%d leave" % (set_index, ep))
        self.barriers.setdefault((set_index, ep), {})[rank] = (arrive, leave)

    def close_barriers(self):
        for ranks in self.barriers.values():
            assert len(ranks) == self.G, "every rank must issue every barrier"
            for (a, _), (_, l) in itertools.product(ranks.values(), ranks.values()):
                self.succ[a].add(l)

    # ---- regions -------------------------------------------------------------------------------------------------
    def owner(self, u):
        return u % self.G

    def phase1(self, r, u, st):
        G = self.G
        w = {("mm", r, u)}
        if self.staged:   # local staging planes; the own block is the slab itself (fill_half_staged)
            w |= {("stage", r, u, g) for g in range(G) if g != r} | {("slab", r, u, r), ("nyqstage", r, u)}
        else:             # peer stores straight into the column owners' slabs (fill_half_peers)
            w |= {("slab", g, u, r) for g in range(G)} | {("nyq", self.owner(u), u, r)}
        self.op(r, st, "phase1 u%d" % u, reads={("in", r)}, writes=w)

    def exchange1(self, r, u, st):
        G = self.G
        self.op(r, st, "exchange1 u%d" % u,
                reads={("stage", r, u, g) for g in range(G) if g != r} | {("nyqstage", r, u)},
                writes={("slab", g, u, r) for g in range(G) if g != r} | {("nyq", self.owner(u), u, r)})

    def phase2(self, r, u, st):
        reg = {("slab", r, u, q) for q in range(self.G)}
        if self.owner(u) == r:
            reg |= {("nyq", r, u, q) for q in range(self.G)}
        self.op(r, st, "phase2 u%d" % u, reads=reg, writes=reg)

    def exchange3(self, r, u, st):
        G = self.G
        reads = {("slab", r, u, q) for q in range(G) if q != r}
        writes = {("stage", q, u, r) for q in range(G) if q != r}
        if self.owner(u) == r:
            reads |= {("nyq", r, u, q) for q in range(G)}
            writes |= {("nyqstage", q, u) for q in range(G)}
        self.op(r, st, "exchange3 u%d" % u, reads=reads, writes=writes)

    def phase3(self, r, u, st):
        G = self.G
        if self.staged:
            reads = {("stage", r, u, g) for g in range(G) if g != r} | {("slab", r, u, r), ("nyqstage", r, u)}
        else:
            reads = {("slab", g, u, r) for g in range(G)} | {("nyq", self.owner(u), u, r)}
        self.op(r, st, "phase3 u%d" % u, reads=reads, writes={("raw", r, u), ("mm", r, u)})
        self.op(r, st, "minmax_decode u%d" % u, reads={("mm", r, u)}, writes={("mmf", r, u)})

    def restore(self, r):
        """fdr_shard_restore_rows, statement by statement (shard.cu, the pipelined driver)."""
        U, staged = self.U, self.staged
        cmp_, link, bar, caller = "cmp", "link", "bar", "caller"
        self.record(r, ("fork",), caller)
        for s in (cmp_, link, bar):
            self.wait(r, ("fork",), s)
        for u in range(U):
            self.phase1(r, u, cmp_)
            self.record(r, (0, u), cmp_)
        for u in range(U):
            if staged:
                self.wait(r, (0, u), link)
                self.exchange1(r, u, link)
                self.record(r, (1, u), link)
                self.wait(r, (1, u), bar)
            else:
                self.wait(r, (0, u), bar)
            self.barrier(r, 2 * u, bar)
            self.record(r, (2, u), bar)
        for u in range(U):
            self.wait(r, (2, u), cmp_)
            self.phase2(r, u, cmp_)
            self.record(r, (3, u), cmp_)
        for u in range(U):
            if staged:
                self.wait(r, (3, u), link)
                self.exchange3(r, u, link)
                self.record(r, (4, u), link)
                self.wait(r, (4, u), bar)
            else:
                self.wait(r, (3, u), bar)
            self.barrier(r, 2 * u + 1, bar)
            self.record(r, (5, u), bar)
        for u in range(U):
            self.wait(r, (5, u), cmp_)
            self.phase3(r, u, cmp_)
            self.record(r, (6, u), cmp_)
        # fdr_shard_minmax_allreduce: one kernel that is also a barrier (set 15) over the [C][2] extrema
        mmf = {("mmf", r, u) for u in range(U)}
        self.op(r, cmp_, "minmax pre", reads=mmf)
        self.barrier(r, 15, cmp_)
        self.op(r, cmp_, "minmax fold", writes=mmf)
        self.op(r, cmp_, "phase4", reads=mmf | {("raw", r, u) for u in range(U)}, writes={("out", r)})
        self.record(r, ("join",), cmp_)
        self.wait(r, ("join",), caller)

    # ---- analysis ------------------------------------------------------------------------------------------------
    def races(self):
        self.close_barriers()
        n = len(self.ops)
        order = list(range(n))   # ops were appended in an order compatible with the edges except barrier cross edges: do a real closure
        reach = [0] * n
        # reverse topological order through DFS
        seen, topo = [False] * n, []
        for root in order:
            if seen[root]:
                continue
            stack = [(root, iter(self.succ[root]))]
            seen[root] = True
            while stack:
                node, it = stack[-1]
                for nx in it:
                    if not seen[nx]:
                        seen[nx] = True
                        stack.append((nx, iter(self.succ[nx])))
                        break
                else:
                    topo.append(node)
                    stack.pop()
        for node in topo:   # successors are finished before their predecessors
            m = 1 << node
            for nx in self.succ[node]:
                m |= reach[nx]
            reach[node] = m
        touch = {}
        for i, (_, _, _, rd, wr) in enumerate(self.ops):
            for reg in rd:
                touch.setdefault(reg, []).append((i, False))
            for reg in wr:
                touch.setdefault(reg, []).append((i, True))
        bad = []
        for reg, acc in touch.items():
            for (a, wa), (b, wb) in itertools.combinations(acc, 2):
                if a == b or not (wa or wb):
                    continue
                if not (reach[a] >> b) & 1 and not (reach[b] >> a) & 1:
                    bad.append((reg, self.ops[a][:3], self.ops[b][:3]))
        return bad


def build(G, U, staged, images=2, drop=None):
    s = Schedule(G, U, staged, drop)
    for _ in range(images):   # the caller issues the images back to back on its stream, every rank the same sequence
        for r in range(G):
            s.op(r, "caller", "write input", writes={("in", r)})   # e.g. the H2D copy of the next image's rows
            s.restore(r)
            s.op(r, "caller", "read output", reads={("out", r)})
    return s


@pytest.mark.parametrize("G", [2, 4, 8])
@pytest.mark.parametrize("staged", [True, False])
def test_native_driver_schedule_has_no_unordered_conflicts(G, staged):
    assert build(G, 3, staged).races() == []


@pytest.mark.parametrize("drop,staged", [
    (("wait", 2, "cmp"), True),    # column phase not waiting for the barrier after exchange 1
    (("wait", 5, "cmp"), True),    # inverse rows not waiting for the barrier after exchange 3
    (("wait", 0, "link"), True),   # exchange 1 not waiting for the rows it sends
    (("wait", 3, "link"), True),   # exchange 3 not waiting for the column phase
    (("wait", 1, "bar"), True),    # barrier signalled before this rank's exchange 1 has landed
    (("wait", 0, "bar"), False),   # fused form: barrier signalled before this rank's scatter
    (("wait", "fork", "cmp"), True),    # the passes not ordered after the caller's stream (next image's input, last image's output)
    (("wait", "join", "caller"), True),  # the caller not waiting for the pack
])
def test_model_detects_a_dropped_wait(drop, staged):
    """Negative controls: the checker is only worth something if it sees the races it is meant to exclude."""
    assert build(4, 3, staged, drop=drop).races() != []


# ---- the pipelined HOST entry point (csrc/capi.cu, fdr_restore_images_host_u8): H2D of chunk k+1, restoration of chunk k and
# D2H of chunk k-1 on three streams over double-buffered device staging; the reference does memcpy -> H2D -> compute -> D2H ->
# sync serially per channel (fft_gpu.cu:347-349, 373-374) ----
def host_pipeline(chunks, drop=None):
    s = Schedule(1, 1, False, drop)
    for k in range(chunks):
        b = k & 1
        if k >= 2:
            s.wait(0, ("cmp", b), "in")           # chunk k-2 no longer reads din[b]
        s.op(0, "in", "H2D %d" % k, reads={("host_in", k)}, writes={("din", b)})
        s.record(0, ("in", b), "in")
        s.wait(0, ("in", b), "cmp")
        if k >= 2:
            s.wait(0, ("out", b), "cmp")          # chunk k-2's D2H has drained dout[b]
        s.op(0, "cmp", "restore %d" % k, reads={("din", b)}, writes={("dout", b), ("workspace",)})
        s.record(0, ("cmp", b), "cmp")
        s.wait(0, ("cmp", b), "out")
        s.op(0, "out", "D2H %d" % k, reads={("dout", b)}, writes={("host_out", k)})
        s.record(0, ("out", b), "out")
    return s


def test_host_pipeline_schedule_has_no_unordered_conflicts():
    for chunks in (1, 2, 3, 7):
        assert host_pipeline(chunks).races() == []


@pytest.mark.parametrize("drop", [("wait", "cmp", "in"), ("wait", "out", "cmp"), ("wait", "in", "cmp"), ("wait", "cmp", "out")])
def test_host_pipeline_model_detects_a_dropped_wait(drop):
    assert host_pipeline(5, drop=drop).races() != []
