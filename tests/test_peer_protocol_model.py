"""CPU model of the cross-rank word protocol of the persistent peer kernel (csrc/chunk.inl, peer mode).

The kernel sends every cosine as ONE 64-bit word -- position in the global batch (18 bits) | step tag (14 bits) |
cosine bits -- into a per-sender inbox slot that is reused every second step, and the reader accepts a word when
its tag matches the step and its position lies inside the global batch.  There is no flag and no fence, so the
argument rests on two invariants, checked here against a brute-force history of what each slot holds:
  * a slot the reader looks at (k < list length of step t) can only hold the word of step t or of step t-2
    (same parity buffer) or the invalid word -- never anything older -- and the tag tells t from t-2;
  * this needs the writer to overwrite the slots that dropped out of the list ([len(t), len(t-2)) of the same
    parity) with the invalid word, every step.
The constants mirror chunk.inl (kPairPosBits, kPairInvalid, pair_tag)."""
import numpy as np

POS_BITS = 18
POS_MASK = (1 << POS_BITS) - 1
TAG_MASK = (1 << (32 - POS_BITS)) - 1
INVALID = (1 << 64) - 1


def pair_tag(t):
    return (t >> 1) & TAG_MASK


def pack(pos, t, cos_bits):
    return ((pos & POS_MASK) | (pair_tag(t) << POS_BITS)) | (cos_bits << 32)


def valid(word, t, gb):
    lo = word & 0xFFFFFFFF
    return (lo >> POS_BITS) == pair_tag(t) and (lo & POS_MASK) < gb


def run(lengths, cap, gb, invalidate=True, t0=1):
    """Writer + reader over the given per-step list lengths; returns the steps at which the reader would have
    accepted a word that was not written at that step."""
    box = [[INVALID] * cap, [INVALID] * cap]          # two parities, initialised invalid (dist.PeerTrainSession)
    written_at = [[None] * cap, [None] * cap]
    prev_len = [0, 0]
    wrong = []
    for s, n in enumerate(lengths):
        t = t0 + s
        par = t & 1
        # reader BEFORE the writer has written anything of step t: nothing may validate
        for k in range(n):
            if valid(box[par][k], t, gb):
                wrong.append((t, k, "early"))
        # writer: the step's entries, then the tail that dropped out of the list
        for k in range(n):
            box[par][k] = pack((7 * k + t) % gb, t, k)
            written_at[par][k] = t
        if invalidate:
            for k in range(n, min(prev_len[par], cap)):
                box[par][k] = INVALID
                written_at[par][k] = None
        prev_len[par] = n
        # reader after the writer: every entry validates and is this step's
        for k in range(n):
            assert valid(box[par][k], t, gb)
            if written_at[par][k] != t:
                wrong.append((t, k, "stale"))
    return wrong


def test_tagged_words_never_validate_stale_data():
    rng = np.random.RandomState(0)
    cap, gb = 64, 50
    for trial in range(200):
        lengths = rng.randint(0, cap + 1, size=rng.randint(4, 60)).tolist()
        assert run(lengths, cap, gb, t0=int(rng.randint(1, 1 << 20))) == []


def test_tag_period_needs_the_invalidation():
    """Without the tail invalidation a slot can keep a word for 2 * 2^14 steps and validate again: the very case the
    invalidation exists for."""
    cap, gb = 8, 8
    period = 2 * (TAG_MASK + 1)
    lengths = [8] + [0] * (period - 1) + [8]           # slot 7 written at t0, not again until t0 + period
    assert run(lengths, cap, gb, invalidate=False) != []
    assert run(lengths, cap, gb, invalidate=True) == []


def test_positions_of_the_largest_global_batch_fit_the_word():
    # AR_PEER_MAX_RANKS * AR_MAX_BATCH positions, and the all-ones field stays free for the invalid word
    assert 8 * 16384 < POS_MASK
    assert not valid(INVALID, 5, 8 * 16384)
    assert not valid(INVALID, (TAG_MASK << 1) | 1, 8 * 16384)   # even when the tag field happens to match


def test_header_words_carry_the_full_step():
    # header word = 32 payload bits | step << 32: a word of any other step never matches
    for t in (1, 2, 255, 1 << 20):
        word = (0xDEADBEEF) | (t << 32)
        assert (word >> 32) == t and (word & 0xFFFFFFFF) == 0xDEADBEEF
        assert (((t + 2) << 32) >> 32) != t
