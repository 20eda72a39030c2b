"""Token <-> note-event codec and onset-F1 matcher (upstream amt/src/utils/
event_codec.py, tokenizer.py, event2note.py, metrics.py [RECALL -- vocabulary layout unverifiable here]).

Vocabulary (596 ids, config.PROVENANCE['vocab_size']): 0 pad, 1 eos, 2 unk, then the event ranges in order
``shift`` (206 steps of 10 ms), ``pitch`` (128), ``velocity`` (2: 0 = off, 1 = on), ``tie`` (1),
``program`` (128), ``drum`` (128).  Decoding is the MT3 state machine: a ``shift`` token is the ABSOLUTE tick offset
from the segment start (206 values x 10 ms cover one 2.048 s segment; the MT3 run-length rule: consecutive shift
tokens add up, any non-shift event resets the running count, so the time after a run of shifts is the run's sum
measured from the segment start -- never cumulative across events), ``program`` / ``velocity`` set the state,
``pitch`` emits a note-on (velocity 1) or note-off
(velocity 0), ``drum`` emits a drum onset; tokens before ``tie`` declare notes tied over from the previous
segment.  CPU-side, not performance relevant; used for the note-level agreement check of the bf16 path."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, List, Sequence, Tuple

import numpy as np

PAD, EOS, UNK = 0, 1, 2
NUM_SPECIAL = 3
RANGES = [("shift", 206), ("pitch", 128), ("velocity", 2), ("tie", 1), ("program", 128), ("drum", 128)]
STEPS_PER_SECOND = 100
VOCAB_SIZE = NUM_SPECIAL + sum(n for _, n in RANGES)          # 596

_OFFSETS = {}
_o = NUM_SPECIAL
for _name, _n in RANGES:
    _OFFSETS[_name] = (_o, _o + _n)
    _o += _n


def encode_event(kind: str, value: int) -> int:
    lo, hi = _OFFSETS[kind]
    if not 0 <= value < hi - lo:
        raise ValueError(f"{kind} value {value} out of range")
    return lo + value


def decode_event(token: int) -> Tuple[str, int]:
    for kind, (lo, hi) in _OFFSETS.items():
        if lo <= token < hi:
            return kind, token - lo
    return ("special", token)


@dataclass(frozen=True)
class Note:
    onset: float
    pitch: int
    program: int = 0
    is_drum: bool = False
    offset: float = float("nan")


def notes_to_tokens(notes: Sequence[Note], max_len: int = 1024) -> List[int]:
    """Encode note ONSETS (and offsets when finite) of one segment, sorted by time (MT3 event order)."""
    ev = []
    for n in notes:
        ev.append((n.onset, 1, n))
        if np.isfinite(n.offset) and not n.is_drum:
            ev.append((n.offset, 0, n))
    ev.sort(key=lambda e: (round(e[0] * STEPS_PER_SECOND), e[1], e[2].program, e[2].pitch))
    out, cur, prog, vel = [encode_event("tie", 0)], 0, None, None
    for t, on, n in ev:
        step = int(round(t * STEPS_PER_SECOND))
        if step > cur:
            # one ABSOLUTE shift per time change (run-length rule: a target beyond 205 ticks is a run of shift
            # tokens that sums to the absolute tick)
            rest = step
            while rest > 0:
                d = min(205, rest)
                out.append(encode_event("shift", d))
                rest -= d
            cur = step
        if n.is_drum:
            out.append(encode_event("drum", n.pitch))
            continue
        if prog != n.program:
            out.append(encode_event("program", n.program))
            prog = n.program
        if vel != on:
            out.append(encode_event("velocity", on))
            vel = on
        out.append(encode_event("pitch", n.pitch))
    out.append(EOS)
    return out[:max_len]


def tokens_to_notes(tokens: Iterable[int], start_time: float = 0.0) -> List[Note]:
    """One segment's token ids -> notes with onsets in seconds (offsets filled when a note-off follows)."""
    notes: List[Note] = []
    active = {}
    cur, run, prog, vel, seen_tie = 0, 0, 0, 1, False
    for tok in tokens:
        tok = int(tok)
        if tok == EOS:
            break
        if tok in (PAD, UNK):
            continue
        kind, v = decode_event(tok)
        if kind == "shift":
            run += v            # shifts of one run add up ...
            cur = run           # ... to the absolute tick from the segment start
            continue
        run = 0                 # any non-shift event ends the run (MT3 decode_events rule)
        if kind == "program":
            prog = v
        elif kind == "velocity":
            vel = v
        elif kind == "tie":
            seen_tie = True
        elif kind == "drum":
            notes.append(Note(start_time + cur / STEPS_PER_SECOND, v, 128, True))
        elif kind == "pitch":
            if not seen_tie:          # tie section: notes carried over from the previous segment, no new onset
                continue
            t = start_time + cur / STEPS_PER_SECOND
            if vel > 0:
                active[(prog, v)] = len(notes)
                notes.append(Note(t, v, prog, False))
            elif (prog, v) in active:
                i = active.pop((prog, v))
                notes[i] = Note(notes[i].onset, v, prog, False, t)
    return notes


def onset_f1(ref: Sequence[Note], est: Sequence[Note], tolerance: float = 0.05) -> Tuple[float, float, float]:
    """mir_eval-style note-onset precision / recall / F1: a match needs equal pitch (and drum flag) and
    |onset difference| <= tolerance; each note is matched at most once (greedy in time order per pitch,
    which is optimal for interval matching on a line)."""
    def by_pitch(ns):
        d = {}
        for n in ns:
            d.setdefault((n.pitch, n.is_drum), []).append(n.onset)
        return {k: sorted(v) for k, v in d.items()}
    R, E = by_pitch(ref), by_pitch(est)
    matched = 0
    for k, r in R.items():
        e = E.get(k, [])
        i = j = 0
        while i < len(r) and j < len(e):
            if abs(r[i] - e[j]) <= tolerance + 1e-9:
                matched += 1
                i += 1
                j += 1
            elif e[j] < r[i]:
                j += 1
            else:
                i += 1
    p = matched / len(est) if est else float(not ref)
    rc = matched / len(ref) if ref else float(not est)
    f = 2 * p * rc / (p + rc) if p + rc > 0 else 0.0
    return p, rc, f


def batch_tokens_to_notes(tokens: np.ndarray, segment_seconds: float = 32767 / 16000.0) -> List[Note]:
    """(n_seg, L) or (n_seg, C, L) token array of consecutive segments -> notes on the global time axis."""
    tokens = np.asarray(tokens)
    out: List[Note] = []
    for s in range(tokens.shape[0]):
        rows = tokens[s].reshape(-1, tokens.shape[-1])
        for row in rows:
            out += tokens_to_notes(row, s * segment_seconds)
    return out
