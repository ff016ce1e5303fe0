"""Model check of the mbarrier protocol of the DRAFT CTA-pair convolution kernel (csrc/experimental/seg_conv_tc2_pair.cuh,
DESIGN.md §9.1) — the part of a tcgen05 pipeline that usually goes wrong (phase parities, arrival counts, who signals
whom) and that cannot be exercised without a GPU.

The model transcribes the kernel's control flow role by role (two TMA producers, the leader's MMA issuer, 2 x 8 epilogue
warps) onto an abstract mbarrier (arrival count, transaction bytes, phase parity; `wait(parity)` passes when the barrier is
no longer in the phase of that parity) and runs it under many random interleavings, with the asynchronous parts — TMA
completions, the in-order tensor pipe, commit arrivals — delivered at arbitrary later times.  Checked on every schedule:
  * no deadlock: every role finishes;
  * no shared-memory hazard: a TMA load never lands in a slot whose previous contents the tensor pipe has not consumed,
    and an MMA only ever reads slots holding BOTH CTAs' tiles of exactly its k-iteration;
  * no accumulator hazard: the MMA never starts a tile in a TMEM buffer that some epilogue warp of either CTA has not
    drained, and an epilogue warp only reads a buffer that holds its tile, complete.
It also checks the tile schedule: every element of the M x N output is stored exactly once, for ragged M and N.
A protocol variant with a deliberately wrong parity must be caught (the model is not vacuous)."""
import random

import pytest

STAGES, EPI_WARPS, BYTES = 4, 8, 32768  # P2_STAGES, epilogue warps per CTA, P2_STAGE_BYTES


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0

    def _maybe_complete(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count

    def arrive(self):
        assert self.pending > 0, "more arrivals than the barrier was initialised for"
        self.pending -= 1
        self._maybe_complete()

    def arrive_expect_tx(self, nbytes):
        self.tx += nbytes
        self.arrive()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        self._maybe_complete()

    def passed(self, parity):
        return (self.phase & 1) != parity


class Model:
    def __init__(self, tiles, iters, rng, bug=None):
        self.tiles, self.iters, self.rng, self.bug = tiles, iters, rng, bug
        self.full = [[MBar(1) for _ in range(STAGES)] for _ in range(2)]      # only CTA 0's is signalled / waited on
        self.empty = [[MBar(1) for _ in range(STAGES)] for _ in range(2)]
        self.tfull = [[MBar(1) for _ in range(2)] for _ in range(2)]
        self.tempty = [[MBar(2 * EPI_WARPS) for _ in range(2)] for _ in range(2)]  # only CTA 0's is used
        self.smem = [[None] * STAGES for _ in range(2)]       # k-iteration whose tiles the slot holds
        self.consumed = [[True] * STAGES for _ in range(2)]   # the tensor pipe is done with the slot's contents
        self.acc_tile = [None, None]                          # tile whose accumulation is complete in the buffer
        self.acc_writing = [None, None]                       # tile currently being accumulated
        self.drained = [[EPI_WARPS, EPI_WARPS], [EPI_WARPS, EPI_WARPS]]  # [cta][buf]: warps that have read the last tile
        self.async_loads = []                                 # TMA copies in flight (complete in any order)
        self.pipe = []                                        # tensor pipe: MMAs and commits, strictly in order

    # ---- roles (generators yield a predicate to wait for, or None for a plain scheduling point) ----
    def producer(self, r):
        it = 0
        for _t in range(self.tiles):
            for _j in range(self.iters):
                st, ph = it % STAGES, (it // STAGES) & 1
                par = (ph ^ 1) if self.bug != "producer_parity" else ph
                yield lambda st=st, par=par: self.empty[r][st].passed(par)
                if r == 0:
                    self.full[0][st].arrive_expect_tx(2 * BYTES if self.bug != "half_tx" else BYTES)
                self.async_loads.append((r, st, it))
                it += 1
                yield None

    def mma(self):
        it = 0
        for i in range(self.tiles):
            b = i & 1
            par = (((i >> 1) & 1) ^ 1) if self.bug != "tempty_parity" else ((i >> 1) & 1)
            yield lambda b=b, par=par: self.tempty[0][b].passed(par)
            for j in range(self.iters):
                st, ph = it % STAGES, (it // STAGES) & 1
                yield lambda st=st, ph=ph: self.full[0][st].passed(ph)
                self.pipe.append(("mma", i, b, st, it, j == 0))
                self.pipe.append(("commit_empty", st))
                it += 1
                yield None
            self.pipe.append(("commit_tfull", i, b))
            yield None

    def epilogue(self, r, w):
        for i in range(self.tiles):
            b = i & 1
            yield lambda b=b, i=i: self.tfull[r][b].passed((i >> 1) & 1)
            assert self.acc_tile[b] == i, f"epilogue of CTA {r} read buffer {b} holding tile {self.acc_tile[b]} instead of {i}"
            yield None                      # slice 0 -> staging -> store
            self.drained[r][b] += 1         # both slices are in registers
            self.tempty[0][b].arrive()      # remote arrive on the leader's barrier (mbar_arrive_cta0)
            yield None

    # ---- asynchronous engines ----
    def step_tma(self):
        k = self.rng.randrange(len(self.async_loads))
        r, st, it = self.async_loads.pop(k)
        assert self.consumed[r][st], f"TMA overwrote slot {st} of CTA {r} before the tensor pipe consumed it"
        self.smem[r][st], self.consumed[r][st] = it, False
        self.full[0][st].complete_tx(BYTES)  # cta_group::2 load: bytes land on the LEADER's barrier from either CTA

    def step_pipe(self):
        op = self.pipe.pop(0)
        if op[0] == "mma":
            _, i, b, st, it, first = op
            assert self.smem[0][st] == it and self.smem[1][st] == it, f"MMA of k-iteration {it} read slots holding {self.smem[0][st]}/{self.smem[1][st]}"
            if first:
                assert self.drained[0][b] == EPI_WARPS and self.drained[1][b] == EPI_WARPS, f"MMA restarted buffer {b} before it was drained"
                self.acc_writing[b], self.acc_tile[b] = i, None
                self.drained[0][b] = self.drained[1][b] = 0
            assert self.acc_writing[b] == i
        elif op[0] == "commit_empty":
            st = op[1]
            for r in (0, 1):                # multicast 0b11
                self.consumed[r][st] = True
                self.empty[r][st].arrive()
        else:
            _, i, b = op
            self.acc_tile[b], self.acc_writing[b] = i, None
            for r in (0, 1):
                self.tfull[r][b].arrive()

    def run(self):
        roles = {("prod", 0): self.producer(0), ("prod", 1): self.producer(1), ("mma",): self.mma()}
        for r in (0, 1):
            for w in range(EPI_WARPS):
                roles[("epi", r, w)] = self.epilogue(r, w)
        waiting = {k: None for k in roles}  # None = runnable
        done = set()
        for _ in range(2_000_000):
            options = [k for k in roles if k not in done and (waiting[k] is None or waiting[k]())]
            if self.async_loads:
                options.append("tma")
            if self.pipe:
                options.append("pipe")
            if not options:
                assert len(done) == len(roles), f"deadlock: stuck roles {sorted(k for k in roles if k not in done)}"
                return
            k = self.rng.choice(options)
            if k == "tma":
                self.step_tma()
            elif k == "pipe":
                self.step_pipe()
            else:
                try:
                    waiting[k] = next(roles[k])
                except StopIteration:
                    done.add(k)
        raise AssertionError("model did not terminate")


@pytest.mark.parametrize("tiles,iters", [(1, 1), (1, 9), (2, 3), (3, 1), (4, 5), (5, 2), (7, 4)])
def test_pair_protocol_has_no_deadlock_or_hazard(tiles, iters):
    for seed in range(200):
        Model(tiles, iters, random.Random(1000 * tiles + 10 * iters + seed)).run()


@pytest.mark.parametrize("bug", ["tempty_parity", "producer_parity", "half_tx"])
def test_model_catches_protocol_mistakes(bug):
    """Mutations the model must notice, otherwise it proves nothing: the MMA warp waiting on tempty with the wrong parity
    (starts overwriting an undrained accumulator), the producers waiting on empty with the wrong parity (deadlock at the
    first k-iteration), the leader expecting only its own CTA's bytes (the MMA reads a slot before the peer's tile landed)."""
    caught = 0
    for seed in range(40):
        try:
            Model(4, 6, random.Random(seed), bug=bug).run()
        except AssertionError:
            caught += 1
    assert caught >= (40 if bug != "half_tx" else 30), caught  # half_tx needs an unlucky interleaving to show


@pytest.mark.parametrize("M,N,clusters", [(256, 256, 1), (1000, 300, 3), (17424, 2048, 74), (513, 64, 2), (130, 257, 5)])
def test_pair_tile_schedule_covers_the_output_exactly_once(M, N, clusters):
    """Transcription of the tile / box arithmetic of conv_gemm_tc2_pair: cluster c handles tiles c, c + n_clusters, ...;
    CTA `rank` owns rows m0 = (t / n_tiles) * 256 + 128 * rank; each (lane group, 64-column slice) box is stored by TMA,
    which clips at the tensor bounds."""
    import numpy as np
    BN = 256
    n_tiles, m_tiles = -(-N // BN), -(-M // 256)
    hit = np.zeros((M, N), dtype=np.int32)
    for c in range(clusters):
        for t in range(c, m_tiles * n_tiles, clusters):
            for rank in (0, 1):
                m0, n0 = (t // n_tiles) * 256 + rank * 128, (t % n_tiles) * BN
                for lg in range(4):
                    for sl in range(4):
                        r0, c0 = m0 + lg * 32, n0 + sl * 64
                        if r0 < M and c0 < N:            # the kernel's guard; TMA clips the rest of the box
                            hit[r0:min(r0 + 32, M), c0:min(c0 + 64, N)] += 1
    assert hit.min() == 1 and hit.max() == 1
