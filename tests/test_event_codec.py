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
