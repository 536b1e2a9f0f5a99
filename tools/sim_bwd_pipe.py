"""Protocol simulator for attn_bwd_kernel<PIPE=true> (bring-up aid, CPU only).

The pipelined attention backward hands shared-memory tiles and TMEM accumulators between four
roles through phase-parity mbarriers.  This script replays the SAME wait / arrive / commit sequence
as the kernel's roles (TMA warp, MMA warp, compute warps, write-out warpgroup) as cooperating
generators under a randomised discrete-event schedule, with mbarriers modelled exactly (a wait on
parity P passes iff the barrier's current phase parity != P, so a waiter that lags two phases
false-passes just like the hardware), and checks at every consumption that the resource holds the
data it is supposed to hold.  It finds deadlocks, parity slips and read-before-write /
write-before-read hazards of the *protocol*; it does not execute the CUDA code.

  python tools/sim_bwd_pipe.py            # all (KT, QT) shapes, 300 random schedules each
"""
import heapq
import itertools
import random
import sys


class MBar:
  def __init__(self, name, count):
    self.name, self.count, self.pending, self.phase = name, count, count, 0

  def arrive(self):
    self.pending -= 1
    assert self.pending >= 0, f"{self.name}: too many arrivals"
    if self.pending == 0:
      self.pending = self.count
      self.phase += 1

  def test(self, parity):
    return (self.phase & 1) != parity


class Sim:
  def __init__(self, KT, QT, items, seed):
    self.KT, self.QT, self.items = KT, QT, items
    self.pairs = KT * QT
    self.rng = random.Random(seed)
    self.now = 0.0
    self.events = []            # (time, seq, callable)
    self.seq = itertools.count()
    B = lambda n, c=1: MBar(n, c)
    self.full_kv = [B("full_kv0"), B("full_kv1")]
    self.empty_kv = [B("empty_kv0"), B("empty_kv1")]
    self.full_q = [B("full_q0"), B("full_q1")]
    self.empty_q = [B("empty_q0"), B("empty_q1")]
    self.sdp_full, self.sdp_empty = B("sdp_full"), B("sdp_empty")
    self.pds_full, self.pds_empty = B("pds_full"), B("pds_empty")
    self.dkv_full, self.dkv_empty = B("dkv_full"), B("dkv_empty")
    self.dq_full, self.dq_empty = B("dq_full"), B("dq_empty")
    # resources: what they currently hold / who still reads them
    self.kv = [None, None]      # (item, kt)
    self.q = [None, None]       # (item, qt)
    self.kv_readers = [0, 0]    # MMAs issued but not yet executed that read the slot
    self.q_readers = [0, 0]
    self.sdp = None             # (item, j) held in TMEM S/dP, None = consumed
    self.pds = None             # (item, j) in the P/dS smem tiles
    self.pds_readers = 0
    self.dkv_acc = None         # (item, kt, number of qt accumulated)
    self.dq_acc = None          # (item, set of (kt, qt) accumulated)
    self.stats = [None, None]   # lse/delta double buffer: item
    self.mma_queue = []         # in-order tensor pipe: list of ops
    self.mma_busy = False
    self.done_items_out = 0
    self.log = []

  # ---------------------------------------------------------------- scheduling helpers
  def at(self, dt, fn):
    heapq.heappush(self.events, (self.now + dt, next(self.seq), fn))

  def delay(self, lo=0.1, hi=2.0):
    return self.rng.uniform(lo, hi)

  def run_role(self, gen):
    """Advance a role generator; it yields ('wait', bar, parity) | ('sleep', dt)."""
    def step():
      try:
        req = next(gen)
      except StopIteration:
        return
      if req[0] == "sleep":
        self.at(req[1], step)
      else:
        _, bar, parity = req
        def poll():
          if bar.test(parity):
            step()
          else:
            self.at(0.05, poll)
        poll()
    step()

  # ---------------------------------------------------------------- tensor pipe (in order, async)
  def mma_issue(self, op):
    self.mma_queue.append(op)
    if not self.mma_busy:
      self.mma_busy = True
      self.at(self.delay(0.2, 1.0), self.mma_exec)

  def mma_exec(self):
    op = self.mma_queue.pop(0)
    op()
    if self.mma_queue:
      self.at(self.delay(0.2, 1.0), self.mma_exec)
    else:
      self.mma_busy = False

  def slot_kv(self, it, kt):
    return (it & 1) if self.KT == 1 else kt

  def slot_q(self, it, qt):
    return (it & 1) if self.QT == 1 else qt

  # ---------------------------------------------------------------- roles
  def tma(self):
    fills_kv, fills_q = [0, 0], [0, 0]
    for it in range(self.items):
      def load_kv(t):
        sl = self.slot_kv(it, t)
        yield ("wait", self.empty_kv[sl], (fills_kv[sl] & 1) ^ 1)
        fills_kv[sl] += 1
        assert self.kv_readers[sl] == 0, f"TMA overwrites K/V slot {sl} still read by MMAs"
        def land(sl=sl, tag=(it, t)):
          self.kv[sl] = tag
          self.full_kv[sl].arrive()
        self.at(self.delay(0.5, 6.0), land)
      def load_q(t):
        sl = self.slot_q(it, t)
        yield ("wait", self.empty_q[sl], (fills_q[sl] & 1) ^ 1)
        fills_q[sl] += 1
        assert self.q_readers[sl] == 0, f"TMA overwrites Q/dO slot {sl} still read by MMAs"
        def land(sl=sl, tag=(it, t)):
          self.q[sl] = tag
          self.full_q[sl].arrive()
        self.at(self.delay(0.5, 6.0), land)
      yield from load_kv(0)
      for t in range(self.QT):
        yield from load_q(t)
      if self.KT > 1:
        yield from load_kv(1)

  def mma(self):
    use_kv, use_q = [0, 0], [0, 0]
    cnt = {"sdp": 0, "grad": 0, "kt": 0}

    def sdp(it, j):
      kt, qt = divmod(j, self.QT)
      skv, sq = self.slot_kv(it, kt), self.slot_q(it, qt)
      if qt == 0:
        yield ("wait", self.full_kv[skv], use_kv[skv] & 1)
      if kt == 0:
        yield ("wait", self.full_q[sq], use_q[sq] & 1)
      yield ("wait", self.sdp_empty, (cnt["sdp"] & 1) ^ 1)
      cnt["sdp"] += 1
      self.kv_readers[skv] += 1
      self.q_readers[sq] += 1
      def ex():
        assert self.kv[skv] == (it, kt), f"S/dP({it},{j}) read K/V slot {skv} holding {self.kv[skv]}"
        assert self.q[sq] == (it, qt), f"S/dP({it},{j}) read Q/dO slot {sq} holding {self.q[sq]}"
        assert self.sdp is None, f"S/dP({it},{j}) overwrites unread S/dP {self.sdp}"
        self.sdp = (it, j)
        self.kv_readers[skv] -= 1
        self.q_readers[sq] -= 1
        self.sdp_full.arrive()
      self.mma_issue(ex)
      yield ("sleep", self.delay(0.05, 0.3))

    def grads(it, j):
      kt, qt = divmod(j, self.QT)
      skv, sq = self.slot_kv(it, kt), self.slot_q(it, qt)
      ph = it & 1
      yield ("wait", self.pds_full, cnt["grad"] & 1)
      cnt["grad"] += 1
      if qt == 0:
        yield ("wait", self.dkv_empty, (cnt["kt"] & 1) ^ 1)
      if kt == 0 and qt == 0:
        yield ("wait", self.dq_empty, ph ^ 1)
      last_q, last_k = qt == self.QT - 1, kt == self.KT - 1
      self.kv_readers[skv] += 1
      self.q_readers[sq] += 1
      self.pds_readers += 1
      def ex():
        assert self.pds == (it, j), f"grads({it},{j}) read P/dS holding {self.pds}"
        assert self.kv[skv] == (it, kt) and self.q[sq] == (it, qt), "grads read a refilled operand tile"
        if qt == 0:
          assert self.dkv_acc is None, f"dV/dK of ({it},{kt}) start over undrained {self.dkv_acc}"
          self.dkv_acc = (it, kt, 0)
        assert self.dkv_acc[:2] == (it, kt)
        self.dkv_acc = (it, kt, self.dkv_acc[2] + 1)
        if kt == 0 and qt == 0:
          assert self.dq_acc is None, f"dQ of item {it} starts over undrained {self.dq_acc}"
          self.dq_acc = (it, set())
        assert self.dq_acc[0] == it
        self.dq_acc[1].add((kt, qt))
        self.kv_readers[skv] -= 1
        self.q_readers[sq] -= 1
        self.pds_readers -= 1
        self.pds_empty.arrive()
        if last_q:
          self.dkv_full.arrive()
          self.empty_kv[skv].arrive()
        if last_k:
          self.empty_q[sq].arrive()
        if last_q and last_k:
          self.dq_full.arrive()
      self.mma_issue(ex)
      if last_q:
        cnt["kt"] += 1
        use_kv[skv] += 1
      if last_k:
        use_q[sq] += 1
      yield ("sleep", self.delay(0.05, 0.3))

    def first_pair_ready(it):
      skv, sq = self.slot_kv(it, 0), self.slot_q(it, 0)
      return self.full_kv[skv].test(use_kv[skv] & 1) and self.full_q[sq].test(use_q[sq] & 1)

    if self.items:
      yield from sdp(0, 0)
    for it in range(self.items):
      for j in range(self.pairs):
        deferred = False
        if j + 1 < self.pairs:
          yield from sdp(it, j + 1)
        elif it + 1 < self.items:
          if first_pair_ready(it + 1):
            yield from sdp(it + 1, 0)
          else:
            deferred = True
        yield from grads(it, j)
        if deferred:
          yield from sdp(it + 1, 0)

  def compute(self):
    pair_cnt = 0
    def prologue(pit):
      self.stats[pit & 1] = pit
    if self.items:
      prologue(0)
    for it in range(self.items):
      yield ("sleep", self.delay(0.01, 0.1))       # named barrier
      if it + 1 < self.items:
        prologue(it + 1)
      for j in range(self.pairs):
        pp = pair_cnt & 1
        yield ("wait", self.sdp_full, pp)
        assert self.sdp == (it, j), f"compute expects S/dP ({it},{j}), TMEM holds {self.sdp}"
        assert self.stats[it & 1] == it, f"statistics buffer of item {it} holds {self.stats[it & 1]}"
        yield ("sleep", self.delay(0.3, 1.5))      # exp
        yield ("wait", self.pds_empty, pp ^ 1)
        assert self.pds_readers == 0, "P/dS rewritten while gradient MMAs still read it"
        self.sdp = None                            # dP read
        self.sdp_empty.arrive()
        yield ("sleep", self.delay(0.2, 1.0))      # write P / dS
        self.pds = (it, j)
        self.pds_full.arrive()
        pair_cnt += 1

  def writeout(self):
    kt_cnt = 0
    for it in range(self.items):
      for kt in range(self.KT):
        yield ("wait", self.dkv_full, kt_cnt & 1)
        assert self.dkv_acc == (it, kt, self.QT), f"write-out expects dV/dK ({it},{kt}) complete, got {self.dkv_acc}"
        yield ("sleep", self.delay(0.2, 1.5))
        self.dkv_acc = None
        self.dkv_empty.arrive()
        kt_cnt += 1
      yield ("wait", self.dq_full, it & 1)
      assert self.dq_acc[0] == it and len(self.dq_acc[1]) == self.pairs, f"dQ of item {it} incomplete: {self.dq_acc}"
      yield ("sleep", self.delay(0.2, 2.0))
      self.dq_acc = None
      self.dq_empty.arrive()
      self.done_items_out += 1

  def run(self):
    for role in (self.tma(), self.mma(), self.compute(), self.writeout()):
      self.run_role(role)
    steps = 0
    while self.events:
      t, _, fn = heapq.heappop(self.events)
      self.now = t
      fn()
      steps += 1
      if steps > 2_000_000:
        raise RuntimeError("deadlock (polling forever)")
    assert self.done_items_out == self.items, f"finished {self.done_items_out}/{self.items} items"


def main():
  n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
  for KT, QT in ((1, 1), (2, 2), (2, 1), (1, 2)):
    for seed in range(n):
      items = 1 + seed % 7
      Sim(KT, QT, items, seed).run()
    print(f"KT={KT} QT={QT}: {n} random schedules, up to 7 items each: OK")


if __name__ == "__main__":
  main()
