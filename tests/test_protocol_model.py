"""A model of the pair kernel's V-ring hand-off (csrc/mla_decode_pair.cu), run under random timing on the CPU.

Not the kernel: a discrete-event restatement of who waits on which mbarrier with which parity -- V producer, TMA landing,
P.V issuer, three softmax groups that publish up to three tiles ahead of P.V -- with `mbarrier.try_wait.parity` modelled
as the hardware defines it (true iff the phase of that parity is not the barrier's current one).  It pins the reasoning
behind the fix of round 2:

* form "softmax": the group that owns a request's last tile zeroes the V rows past the end after a parity wait on v_land
  for units 2g and 2g+1.  The model finds schedules in which that wait returns on the slot's OLDER phase and the zeros
  are written while the slot belongs to unit u-3 (still in flight) -- the bug the timing-perturbation test found on the
  GPU.
* form "two_step": waiting for the phase in front first cannot work -- if unit u has already landed the first wait is
  a wait for a phase that never comes (deadlock).
* form "producer" (shipped): the unit lands on an auxiliary barrier, the producer -- which sees every unit of its ring in
  order -- zeroes, then completes v_land.  No schedule misplaces a zero-fill or deadlocks.
"""
import heapq
import random

import pytest

K_V_UNITS = 3
K_GROUPS = 3


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0

    def _maybe_complete(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count

    def arrive(self, expect_tx=0):
        self.tx += expect_tx
        self.pending -= 1
        self._maybe_complete()

    def complete_tx(self, n):
        self.tx -= n
        self._maybe_complete()

    def try_wait_parity(self, parity):
        # true iff the phase with this parity is not the current (incomplete) one, i.e. is taken to be the preceding,
        # completed one -- which is wrong when the waiter is really two phases ahead
        return (self.phase & 1) != parity


class Sim:
    def __init__(self, seed, n_tiles, form):
        self.rng = random.Random(seed)
        self.t, self.q, self.n = 0.0, [], 0
        self.n_tiles, self.form = n_tiles, form
        self.v_land = [MBar(1) for _ in range(K_V_UNITS)]
        self.v_empty = [MBar(1) for _ in range(K_V_UNITS)]
        self.v_aux = [MBar(1) for _ in range(2)]
        self.p_full = [MBar(1) for _ in range(K_GROUPS)]
        self.p_free = [MBar(1) for _ in range(K_GROUPS)]
        self.owner = [None] * K_V_UNITS       # unit whose bytes the slot holds or is receiving
        self.in_flight = [False] * K_V_UNITS  # a TMA is writing the slot
        self.zeroed = set()                   # units whose tail rows were zeroed while the slot was theirs and landed
        self.misplaced = []                   # (unit meant, unit hit, in flight?)
        self.done = {"producer": False, "issuer": False}
        self.blocked = {}

    def after(self, dt, fn):
        self.n += 1
        heapq.heappush(self.q, (self.t + dt, self.n, fn))

    def spawn(self, name, gen):
        def step(_=None):
            try:
                req = next(gen)
            except StopIteration:
                self.blocked.pop(name, None)
                return
            if req[0] == "sleep":
                self.blocked.pop(name, None)
                self.after(req[1], step)
            else:  # ("wait", bar, parity): poll
                _, bar, parity = req
                if bar.try_wait_parity(parity):
                    self.blocked.pop(name, None)
                    self.after(0.01, step)
                else:
                    self.blocked[name] = req

                    def poll():
                        if bar.try_wait_parity(parity):
                            self.blocked.pop(name, None)
                            step()
                        else:
                            self.after(0.05, poll)
                    self.after(0.05, poll)
        self.after(0.0, step)

    def jitter(self, lo, hi, p_slow=0.0, slow=0.0):
        d = self.rng.uniform(lo, hi)
        if self.rng.random() < p_slow:
            d += self.rng.uniform(0, slow)
        return d

    # ---- actors ----
    def zero_fill(self, u, slot):
        if self.owner[slot] == u and not self.in_flight[slot]:
            self.zeroed.add(u)
        else:
            self.misplaced.append((u, self.owner[slot], self.in_flight[slot]))

    def producer(self):
        last = self.n_tiles - 1
        for g in range(self.n_tiles):
            yield ("sleep", self.jitter(0.0, 0.2, 0.25, 4.0))  # the delays of the perturbation test's V-producer sites
            tail = g == last and self.form == "producer"
            for j in range(2):
                u = 2 * g + j
                slot = u % K_V_UNITS
                yield ("wait", self.v_empty[slot], ((u // K_V_UNITS) & 1) ^ 1)
                yield ("sleep", self.jitter(0.0, 0.1, 0.25, 4.0))
                bar = self.v_aux[j] if tail else self.v_land[slot]
                self.owner[slot], self.in_flight[slot] = u, True
                bar.arrive(expect_tx=1)

                def landed(bar=bar, slot=slot):
                    self.in_flight[slot] = False
                    bar.complete_tx(1)
                self.after(self.jitter(0.6, 1.6), landed)
            if tail:
                for j in range(2):
                    u = 2 * g + j
                    slot = u % K_V_UNITS
                    yield ("wait", self.v_aux[j], 0)
                    self.zero_fill(u, slot)
                    self.v_land[slot].arrive()
        self.done["producer"] = True

    def issuer(self):
        for g in range(self.n_tiles):
            pp = g % K_GROUPS
            yield ("wait", self.p_full[pp], (g // K_GROUPS) & 1)
            for j in range(2):
                u = 2 * g + j
                slot = u % K_V_UNITS
                yield ("wait", self.v_land[slot], (u // K_V_UNITS) & 1)
                assert self.owner[slot] == u and not self.in_flight[slot]
                yield ("sleep", self.jitter(0.4, 0.6))  # the unit's P.V products
                self.v_empty[slot].arrive()
            self.p_free[pp].arrive()
        self.done["issuer"] = True

    def softmax(self, grp):
        last = self.n_tiles - 1
        for g in range(grp, self.n_tiles, K_GROUPS):
            yield ("sleep", self.jitter(0.3, 1.2))  # max pass
            yield ("wait", self.p_free[grp], ((g // K_GROUPS) & 1) ^ 1)  # P pair read by P.V of tile g - 3
            yield ("sleep", self.jitter(0.5, 1.5))  # weighted pass
            if g == last and self.form in ("softmax", "two_step"):
                for j in range(2):
                    u = 2 * g + j
                    slot, ph = u % K_V_UNITS, u // K_V_UNITS
                    if self.form == "two_step" and ph > 0:
                        yield ("wait", self.v_land[slot], (ph - 1) & 1)
                    yield ("wait", self.v_land[slot], ph & 1)
                    self.zero_fill(u, slot)
            self.p_full[grp].arrive()

    def run(self, t_max=400.0):
        self.spawn("producer", self.producer())
        self.spawn("issuer", self.issuer())
        for grp in range(K_GROUPS):
            self.spawn("softmax%d" % grp, self.softmax(grp))
        while self.q and self.t < t_max:
            self.t, _, fn = heapq.heappop(self.q)
            fn()
        return all(self.done.values())


def sweep(form, seeds=300):
    finished, misplaced, zeroed_ok = 0, [], 0
    for seed in range(seeds):
        n_tiles = 6 + seed % 7
        s = Sim(seed, n_tiles, form)
        ok = s.run()
        finished += ok
        misplaced += s.misplaced
        last = n_tiles - 1
        zeroed_ok += ok and {2 * last, 2 * last + 1} <= s.zeroed
    return finished, misplaced, zeroed_ok


def test_parity_semantics():
    b = MBar(1)
    assert not b.try_wait_parity(0) and b.try_wait_parity(1)  # fresh barrier: "the phase before the first" is complete
    b.arrive()
    assert b.try_wait_parity(0) and not b.try_wait_parity(1)
    b.arrive()
    assert b.try_wait_parity(1) and not b.try_wait_parity(0)  # a waiter for phase 0 that comes this late would hang


def test_softmax_side_zero_fill_aliases_under_a_slow_v_producer():
    finished, misplaced, _ = sweep("softmax")
    assert finished == 300  # (the bug never hangs: it passes the wait too early)
    assert misplaced, "the model should reproduce the aliased wait"
    # every misplaced zero-fill found its ring slot still owned by an older unit: u - 6 (consumed, the slot idle until
    # the next TMA overwrites it: harmless) or u - 3 with its TMA in flight -- the corrupting case: zeros under a landing
    # unit of a tile P.V has not reached (on the GPU: one CTA's 128 columns of a d-half wrong for all q rows)
    assert all(meant - hit in (3, 6) for meant, hit, _ in misplaced)
    assert all((meant - hit == 3) == in_flight for meant, hit, in_flight in misplaced)
    assert any(in_flight for _, _, in_flight in misplaced)


def test_two_step_wait_deadlocks():
    finished, _, _ = sweep("two_step", seeds=100)
    assert finished < 100  # a unit that has already landed turns the wait for the phase in front into a wait for the next


@pytest.mark.parametrize("form", ["producer"])
def test_producer_side_zero_fill_is_timing_independent(form):
    finished, misplaced, zeroed_ok = sweep(form)
    assert finished == 300 and not misplaced and zeroed_ok == 300


# ---- the row reference's two lanes ("Deferred rescale") -------------------------------------------
def _move_outcome(rng, read_from_shared, deferred):
    """One reference move of one q row: its two threads (the lanes of the tile's two pages, in different warps) leave the
    group barrier, take r_old, compute the same r_new, and the page-0 thread stores it.  Returns the rescale factors the
    two threads apply (they must agree)."""
    st_ref, r_new = 1.0, 5.0                     # shared row state; both threads have synced their view: rloc == st_ref
    view = {0: st_ref, 1: st_ref}
    t_read = {0: rng.uniform(0, 60), 1: rng.uniform(0, 60)}              # cycles after the barrier: warps drift
    # page 0's store follows its read: behind the drain + rescale of O (~1000 cycles) in the undeferred order, at once
    # in the deferred one
    t_store = t_read[0] + (10 if deferred else 1000)
    fac = {}
    for page in (0, 1):
        r_old = (r_new if (read_from_shared and page == 1 and t_read[1] > t_store) else st_ref) if read_from_shared else view[page]
        fac[page] = 2.0 ** (r_old - (r_new if r_new > r_old else r_old))
    return fac[0], fac[1]


def test_row_reference_is_taken_from_the_threads_own_view():
    rng = random.Random(0)
    # re-reading shared memory: the deferred order lets the partner's store overtake the read in some schedules (the
    # partner then sees "no move" and skips its half of the row); the undeferred order hid it behind the rescale
    bad_deferred = sum(a != b for a, b in (_move_outcome(rng, True, True) for _ in range(2000)))
    bad_undeferred = sum(a != b for a, b in (_move_outcome(rng, True, False) for _ in range(2000)))
    assert bad_deferred > 0 and bad_undeferred == 0
    # the shipped form (both instantiations): r_old is the thread's own synced view
    assert all(a == b for a, b in (_move_outcome(rng, False, d) for d in (True, False) for _ in range(2000)))
