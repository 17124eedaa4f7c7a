"""The algorithm of K1's self-synchronising flavour (csrc/rtj_scan_sync.cu), restated in Python and checked on the CPU
against the oracle's serial walker (oracle.walk_payload, itself pinned to lib/RTjpeg.c:157-186, 2701-2745):

* the 64-state machine (r = places of the block still to fill; r <= 0: the next byte starts a block; a byte 64..127 fills
  byte - 63 places, any other byte one; 0xFF where a block starts is a block of its own) finds exactly the walker's blocks;
* chunks walked from a GUESSED state ("a block starts at the first byte of the lead-in") and then repaired -- a chunk entered
  in another state than its left neighbour ended in is walked again from that state until the new walk falls in step with
  the old one, or to the chunk's end; repeated until nothing changes -- end in the same bit map as one serial walk, whatever
  the stream and however short the lead-in.  Streams that synchronise need one round; a stream of 64-byte blocks needs one
  round per chunk (the case the kernel hands over to the chunk-parallel scan), and is still exact.

No GPU involved: this pins the argument, the -m gpu suite pins the kernel."""
import numpy as np
import pytest

from oracle import oracle as O
from streams import clip


def step(r, b):
    """One byte of the state machine: (state before the byte, byte) -> (state after, does the byte start a block)."""
    if r <= 0:
        return (0 if b == 0xFF else 63), True
    return r - ((b - 63) if 64 <= b <= 127 else 1), False


def serial_starts(pay):
    r, out = 0, np.zeros(len(pay), dtype=bool)
    for i, b in enumerate(pay):
        r, out[i] = step(r, int(b))
    return out


def chunked_starts(pay, nchunks, lead):
    """The kernel's scheme on one segment: returns (bit map, repair rounds, chunks that were entered in a wrong state)."""
    n = len(pay)
    C = max(1, -(-n // nchunks))
    bounds = [(c * C, min((c + 1) * C, n)) for c in range(nchunks) if c * C < n]
    bits = np.zeros(n, dtype=bool)
    entered, left = [], []
    for c, (a, e) in enumerate(bounds):
        r = 0
        if c > 0:
            for i in range(max(a - lead, 0), a):            # lead-in from a guessed state
                r, _ = step(r, int(pay[i]))
        entered.append(r)
        for i in range(a, e):
            r, bits[i] = step(r, int(pay[i]))
        left.append(r)
    rounds, wrong = 0, 0
    while True:
        dirty = [c for c in range(1, len(bounds)) if max(left[c - 1], 0) != max(entered[c], 0)]
        if not dirty:
            break
        rounds += 1
        wrong += len(dirty) if rounds == 1 else 0
        want = {c: left[c - 1] for c in dirty}              # every lane reads its neighbour before anybody repairs
        for c in dirty:
            a, e = bounds[c]
            r = entered[c] = want[c]
            in_step = False
            for i in range(a, e):
                before = bits[i]
                r, bits[i] = step(r, int(pay[i]))
                if bits[i] and before:                      # both walks start a block here: the same walk from now on
                    in_step = True
                    break
            if not in_step:
                left[c] = r
        assert rounds <= len(bounds), "after round k the first k chunks are final"
    return bits, rounds, wrong


def payload_of(s, o, f):
    n = int(O.packet_sizes(s, o)[f])
    return np.asarray(s[int(o[f]) + 12:int(o[f]) + n])


@pytest.mark.parametrize("kw", [dict(Q=128), dict(Q=32), dict(Q=170, noise_y=10, noise_c=3),
                                dict(Q=128, key_rate=4, lm=3, cm=3)])
def test_state_machine_finds_the_walkers_blocks(kw):
    kw = dict(kw)
    Q = kw.pop("Q")
    w, h = 160, 96
    s, o = clip(w, h, Q, 3, **kw)
    nmb = (w // 16) * (h // 16)
    for f in range(3):
        pay = payload_of(s, o, f)
        n, offs, eob = O.walk_payload(pay, nmb, 0, 0)
        assert n == len(pay)
        at = np.flatnonzero(serial_starts(pay))
        coded = np.asarray(eob) > 0                          # the walker gives no offset for a skipped block
        assert len(at) == len(coded)
        assert np.array_equal(at[coded], np.asarray(offs, dtype=np.int64)[coded])
        assert (pay[at[~coded]] == 0xFF).all()


@pytest.mark.parametrize("lead", [0, 16, 64, 256])
@pytest.mark.parametrize("kw", [dict(Q=128), dict(Q=150, key_rate=5, lm=2, cm=2, noise_y=12)])
def test_guess_and_repair_equals_the_serial_walk(kw, lead):
    kw = dict(kw)
    Q = kw.pop("Q")
    s, o = clip(352, 288, Q, 2, **kw)
    for f in range(2):
        pay = payload_of(s, o, f)
        want = serial_starts(pay)
        got, rounds, wrong = chunked_starts(pay, 64, lead)
        assert np.array_equal(got, want)
        assert rounds <= 3                                   # ordinary material: wrongly entered chunks fall in step at once
        if lead >= 256:
            assert wrong <= 16                               # of 64: a long lead-in makes most guesses right


def test_a_stream_that_never_synchronises_is_still_exact():
    """Every block DC + 63 coefficient bytes: a walk entered at the wrong byte stays wrong for ever, the truth moves one
    chunk a round.  (The kernel hands such frames to the chunk-parallel scan; forced, it does what this does.)"""
    rng = np.random.default_rng(3)
    blocks = []
    for _ in range(200):
        b = rng.integers(0, 64, size=64, dtype=np.int64)    # no run tokens, no 0xFF
        b[0] = rng.integers(0, 255)
        blocks.append(b)
    pay = np.concatenate(blocks).astype(np.uint8)
    want = serial_starts(pay)
    assert np.array_equal(np.flatnonzero(want), np.arange(0, len(pay), 64))
    got, rounds, wrong = chunked_starts(pay, 50, 64)        # chunks of 256 bytes, lead-in 64: guesses are right by accident
    assert np.array_equal(got, want)
    got, rounds, wrong = chunked_starts(pay, 47, 48)        # chunks of 273 bytes: every guess is wrong
    assert np.array_equal(got, want)
    assert rounds >= 40 and wrong >= 40


def test_random_bytes_and_damaged_streams():
    """The machine and the repair make no assumption about the stream: random bytes, runs that overshoot, a tail of skip markers."""
    rng = np.random.default_rng(11)
    for trial in range(6):
        pay = rng.integers(0, 256, size=3000, dtype=np.int64).astype(np.uint8)
        if trial % 2:
            pay[2000:] = 0xFF
        want = serial_starts(pay)
        for lead in (0, 32):
            got, _, _ = chunked_starts(pay, 17, lead)
            assert np.array_equal(got, want)


# ---- frames with a raw prefix (quality above 170): the state also holds the block's place in its unit -------------------

def step_raw(r, c, b, thr, unit):
    """(state, place of the block under way, byte) -> (state, place, does the byte start a block).  Inside the prefix
    (r above the plane's threshold 63 - bt8) every byte is a coefficient: one place, whatever its value."""
    if r <= 0:
        return (0 if b == 0xFF else 63), (c + 1) % unit, True
    if r > thr[c]:
        return r - 1, c, False
    return r - ((b - 63) if 64 <= b <= 127 else 1), c, False


def serial_starts_raw(pay, thr, unit):
    r, c, out = 0, unit - 1, np.zeros(len(pay), dtype=bool)
    for i, b in enumerate(pay):
        r, c, out[i] = step_raw(r, c, int(b), thr, unit)
    return out


def chunked_starts_raw(pay, nchunks, lead, thr, unit):
    """The kernel's scheme with the place in the state: a repair walk is in step with the walk made before at a byte where
    both start a block AND both take it for the same block of its unit -- the place of the earlier walk is its place at the
    chunk's first byte plus the starts it made since (the bit map holds no places)."""
    n = len(pay)
    C = max(1, -(-n // nchunks))
    bounds = [(c * C, min((c + 1) * C, n)) for c in range(nchunks) if c * C < n]
    bits = np.zeros(n, dtype=bool)
    entered, left = [], []
    for k, (a, e) in enumerate(bounds):
        st = (0, unit - 1)
        if k > 0:
            for i in range(max(a - lead, 0), a):
                st = step_raw(*st, int(pay[i]), thr, unit)[:2]
        entered.append(st)
        r, c = st
        for i in range(a, e):
            r, c, bits[i] = step_raw(r, c, int(pay[i]), thr, unit)
        left.append((r, c))
    same = lambda x, y: max(x[0], 0) == max(y[0], 0) and x[1] == y[1]
    rounds, wrong = 0, 0
    while True:
        dirty = [k for k in range(1, len(bounds)) if not same(left[k - 1], entered[k])]
        if not dirty:
            break
        rounds += 1
        wrong += len(dirty) if rounds == 1 else 0
        want = {k: left[k - 1] for k in dirty}
        for k in dirty:
            a, e = bounds[k]
            was_c = entered[k][1]
            r, c = entered[k] = want[k]
            ahead = (was_c - c) % unit                      # the earlier walk's place less this walk's
            in_step = False
            for i in range(a, e):
                before = bits[i]
                r, c, bits[i] = step_raw(r, c, int(pay[i]), thr, unit)
                ahead = (ahead + int(before) - int(bits[i])) % unit
                if bits[i] and before and ahead == 0:
                    in_step = True
                    break
            if not in_step:
                left[k] = (r, c)
        assert rounds <= len(bounds)
    return bits, rounds, wrong


def _bt8(Q):
    t = O.tables_from_quality(Q)
    return int(t.lb8), int(t.cb8)


@pytest.mark.parametrize("Q", [171, 200, 255])
def test_raw_prefix_machine_finds_the_walkers_blocks(Q):
    w, h = 160, 96
    s, o = clip(w, h, Q, 2, noise_y=6, noise_c=2)
    lb8, cb8 = _bt8(Q)
    thr = [63 - lb8] * 4 + [63 - cb8] * 2
    nmb = (w // 16) * (h // 16)
    for f in range(2):
        pay = payload_of(s, o, f)
        n, offs, eob = O.walk_payload(pay, nmb, lb8, cb8)
        assert n == len(pay)
        at = np.flatnonzero(serial_starts_raw(pay, thr, 6))
        coded = np.asarray(eob) > 0
        assert len(at) == len(coded)
        assert np.array_equal(at[coded], np.asarray(offs, dtype=np.int64)[coded])


@pytest.mark.parametrize("lead", [0, 64, 192])
@pytest.mark.parametrize("Q", [180, 255])
def test_raw_prefix_guess_and_repair_equals_the_serial_walk(Q, lead):
    s, o = clip(352, 288, Q, 1, noise_y=4, noise_c=1)
    lb8, cb8 = _bt8(Q)
    thr = [63 - lb8] * 4 + [63 - cb8] * 2
    pay = payload_of(s, o, 0)
    want = serial_starts_raw(pay, thr, 6)
    got, rounds, wrong = chunked_starts_raw(pay, 64, lead, thr, 6)
    assert np.array_equal(got, want)
    if lead >= 192:
        assert wrong <= 8 and rounds <= 3                    # a walk that is wrong about byte or place is right ~80 bytes on


def test_raw_prefix_streams_in_step_by_byte_but_not_by_place():
    """Skip markers only: every byte starts a block, so any two walks share every start -- and are the same walk only if
    they also agree on the place.  A repair that looked at the bit map alone would stop at once and keep a wrong exit
    place; with a DC-plus-prefix luma block behind the markers the parse would then go wrong."""
    rng = np.random.default_rng(5)
    unit, thr = 6, [63 - 9] * 4 + [63] * 2
    blocks = []
    for k in range(600):
        if rng.random() < 0.7:
            blocks.append(np.array([0xFF], dtype=np.uint8))
        elif k % 6 < 4:
            blocks.append(np.concatenate([rng.integers(0, 255, size=10), [127]]).astype(np.uint8))
        else:
            blocks.append(np.array([rng.integers(0, 255), 126], dtype=np.uint8))
    pay = np.concatenate(blocks)
    want = serial_starts_raw(pay, thr, unit)
    assert int(want.sum()) == 600
    for lead in (0, 16, 64):
        got, _, _ = chunked_starts_raw(pay, 23, lead, thr, unit)
        assert np.array_equal(got, want)
