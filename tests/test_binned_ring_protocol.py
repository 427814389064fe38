"""Model check of the TMA ring protocol of k_obs_b1_binned_tma (csrc/obs_binned.cuh).

The kernel cannot run here (no GPU), and a wrong slot / phase would hang it, so its control flow -- prologue of
BIN_STAGES - 1 bulk copies per task, one refill per iteration into the slot freed by the previous iteration, wait on
parity (k / BIN_STAGES) & 1 of slot k % BIN_STAGES with k counting stages over ALL tasks of the warp, partial last
stage -- is replayed against a model of the mbarrier semantics (one arrival + transaction bytes complete a phase;
try_wait.parity(p) succeeds iff the most recently completed phase has parity p) for random task sequences.
Asserted: no wait on a phase that was never armed (deadlock), no re-arm of a barrier whose phase is still pending, no
overwrite of a slot that has not been consumed, and every iteration consumes exactly the groups it should."""
import random
import re
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "variational-gridded-gaussian-processes_b200", "csrc", "obs_binned.cuh")).read()
STAGES = int(re.search(r"constexpr int BIN_STAGES = (\d+);", SRC).group(1))
GPS = int(re.search(r"constexpr int BIN_GPS = (\d+);", SRC).group(1))


class Barrier:
    def __init__(self):
        self.phase = 0          # phase in progress
        self.armed = False      # expect_tx issued for the phase in progress, copy not yet landed

    def arm(self):
        assert not self.armed, "re-armed while the previous copy is still pending"
        self.armed = True

    def land(self):             # the bulk copy completes: arrival count and tx bytes reach zero
        assert self.armed
        self.armed = False
        self.phase += 1

    def try_wait(self, parity):  # completed phase = phase - 1
        return self.phase > 0 and ((self.phase - 1) & 1) == parity


def run_warp(task_groups, rng):
    bars = [Barrier() for _ in range(STAGES)]
    slot_content = [None] * STAGES          # (task, first group, n groups) currently in the slot, None = free
    in_flight = []                          # (slot, payload) copies issued but not landed
    k0 = 0
    consumed = []

    def issue(slot, payload):
        assert slot_content[slot] is None, "slot overwritten before it was consumed"
        bars[slot].arm()
        in_flight.append((slot, payload))

    def land_some(force_slot=None):
        # copies land in any order and at any time; force the one we are about to wait for
        rng.shuffle(in_flight)
        keep = []
        for slot, payload in in_flight:
            if slot == force_slot or rng.random() < 0.5:
                bars[slot].land()
                slot_content[slot] = payload
            else:
                keep.append((slot, payload))
        in_flight[:] = keep

    for ti, groups in enumerate(task_groups):
        nst = (groups + GPS - 1) // GPS
        for i in range(STAGES - 1):                      # prologue
            if i < nst:
                issue((k0 + i) % STAGES, (ti, i * GPS, min(GPS, groups - i * GPS)))
        for si in range(nst):
            k = k0 + si
            st = k % STAGES
            nx = si + STAGES - 1
            if nx < nst:                                 # refill the slot freed by the previous iteration
                issue((k + STAGES - 1) % STAGES, (ti, nx * GPS, min(GPS, groups - nx * GPS)))
            parity = (k // STAGES) & 1
            assert bars[st].armed or bars[st].try_wait(parity), "waiting on a phase nobody armed: deadlock"
            land_some(force_slot=st)
            assert bars[st].try_wait(parity), "wrong parity"
            payload = slot_content[st]
            assert payload == (ti, si * GPS, min(GPS, groups - si * GPS)), (payload, ti, si)
            consumed.append(payload)
            slot_content[st] = None                      # __syncwarp: all lanes hold the stage in registers
        k0 += nst
    assert not in_flight
    return consumed


def test_ring_protocol_random_task_sequences():
    rng = random.Random(0)
    for trial in range(300):
        tasks = [rng.choice([1, 1, 2, 3, 4, 5, 7, 8, 16, 63, 64]) for _ in range(rng.randint(1, 12))]
        consumed = run_warp(tasks, rng)
        want = [(ti, g, min(GPS, groups - g)) for ti, groups in enumerate(tasks) for g in range(0, groups, GPS)]
        assert consumed == want


def test_kernel_source_matches_the_modelled_protocol():
    """The expressions the model replays, as they appear in the kernel (guards against drift)."""
    for needle in ["(k0 + i) % BIN_STAGES", "(k + BIN_STAGES - 1) % BIN_STAGES", "(k / BIN_STAGES) & 1u",
                   "k0 += (unsigned int)nst", "const int nx = si + BIN_STAGES - 1", "if (lane == 0 && nx < nst)",
                   "for (int i = 0; i < BIN_STAGES - 1; ++i)", "min(BIN_GPS, t.groups - nx * BIN_GPS)",
                   "min(BIN_GPS, t.groups - i * BIN_GPS)"]:
        assert needle in SRC, needle
