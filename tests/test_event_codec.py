"""Token <-> note-event codec and onset-F1 matcher (CPU)."""
import numpy as np

from yourmt3_b200 import event_codec as EC


def test_vocab_layout():
    assert EC.VOCAB_SIZE == 596
    assert EC.decode_event(EC.encode_event("shift", 205)) == ("shift", 205)
    assert EC.decode_event(EC.encode_event("drum", 127)) == ("drum", 127)
    assert EC.encode_event("drum", 127) == 595


def test_round_trip_and_f1():
    rng = np.random.default_rng(0)
    notes = sorted((EC.Note(round(float(rng.uniform(0, 2.0)), 2), int(rng.integers(36, 96)), int(rng.integers(0, 4)), False,
                            2.04) for _ in range(40)), key=lambda n: n.onset)
    notes += [EC.Note(0.5, 38, 128, True), EC.Note(1.0, 42, 128, True)]
    toks = EC.notes_to_tokens(notes)
    assert toks[-1] == EC.EOS and max(toks) < EC.VOCAB_SIZE
    back = EC.tokens_to_notes(toks + [EC.PAD] * 5)
    p, r, f = EC.onset_f1(notes, back)
    assert f == 1.0 and len(back) == len(notes)
    # drop / shift some notes -> F1 decreases accordingly
    est = [EC.Note(n.onset + (0.2 if i % 4 == 0 else 0.01), n.pitch, n.program, n.is_drum) for i, n in enumerate(notes)]
    p, r, f = EC.onset_f1(notes, est)
    assert 0.6 < f < 0.85
    assert EC.onset_f1([], []) == (1.0, 1.0, 1.0)


def test_tie_section_and_eos():
    toks = [EC.encode_event("program", 3), EC.encode_event("pitch", 60), EC.encode_event("tie", 0),
            EC.encode_event("shift", 10), EC.encode_event("velocity", 1), EC.encode_event("pitch", 64), EC.EOS,
            EC.encode_event("pitch", 70)]
    notes = EC.tokens_to_notes(toks)
    assert len(notes) == 1 and notes[0].pitch == 64 and abs(notes[0].onset - 0.1) < 1e-9 and notes[0].program == 3


def test_shift_tokens_are_absolute_from_segment_start():
    """Hand-written token fixture (NOT produced by notes_to_tokens): MT3 run-length rule.  A shift token carries the
    absolute 10 ms tick from the segment start; consecutive shifts add up; any other event resets the run, so
    `shift 50, pitch, shift 120, pitch` puts the second onset at 1.20 s, not at 1.70 s."""
    sh, pi, ve, tie = (lambda v: EC.encode_event("shift", v)), (lambda v: EC.encode_event("pitch", v)), \
        (lambda v: EC.encode_event("velocity", v)), EC.encode_event("tie", 0)
    toks = [tie, sh(50), ve(1), pi(60), sh(120), pi(62), sh(205), sh(3), pi(64), sh(204), ve(0), pi(60), EC.EOS]
    notes = EC.tokens_to_notes(toks, start_time=10.0)
    assert [(n.pitch, round(n.onset - 10.0, 2)) for n in notes] == [(60, 0.5), (62, 1.2), (64, 2.08)]
    assert abs(notes[0].offset - 12.04) < 1e-9                      # note-off of pitch 60 at absolute tick 204
    # the encoder emits exactly one absolute shift per time change
    enc = EC.notes_to_tokens([EC.Note(0.5, 60), EC.Note(1.2, 62)])
    shifts = [EC.decode_event(t)[1] for t in enc if EC.decode_event(t)[0] == "shift"]
    assert shifts == [50, 120]
