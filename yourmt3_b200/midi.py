"""Standard MIDI File writer for transcribed notes (upstream writes MIDI through mido /
pretty_midi [RECALL], which are not installed here - this is a dependency-free format-1 writer plus the small reader
the tests use to check it).  One track per (program, is_drum); drums go to channel 10 (index 9); tempo 120 bpm,
480 ticks per quarter note, so 1 tick = 1/960 s.  Notes without a finite offset get ``default_duration``."""
from __future__ import annotations

import math
import struct
from typing import Dict, Iterable, List, Sequence, Tuple

from .event_codec import Note

TICKS_PER_QUARTER = 480
TEMPO_US_PER_QUARTER = 500000          # 120 bpm
TICKS_PER_SECOND = TICKS_PER_QUARTER * 1_000_000 // TEMPO_US_PER_QUARTER   # 960


def _vlq(n: int) -> bytes:
    """MIDI variable-length quantity."""
    if n < 0:
        raise ValueError("negative delta time")
    out = [n & 0x7F]
    n >>= 7
    while n:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    return bytes(reversed(out))


def _track(events: List[Tuple[int, int, bytes]]) -> bytes:
    """events: (tick, order, message bytes) -> MTrk chunk (delta-time encoded, end-of-track appended)."""
    events.sort(key=lambda e: (e[0], e[1]))
    body, cur = bytearray(), 0
    for tick, _, msg in events:
        body += _vlq(tick - cur) + msg
        cur = tick
    body += b"\x00\xff\x2f\x00"
    return b"MTrk" + struct.pack(">I", len(body)) + bytes(body)


def notes_to_midi_bytes(notes: Sequence[Note], default_duration: float = 0.1, velocity: int = 100) -> bytes:
    groups: Dict[Tuple[int, bool], List[Note]] = {}
    for n in notes:
        groups.setdefault((0 if n.is_drum else int(n.program), bool(n.is_drum)), []).append(n)
    tracks = [_track([(0, 0, b"\xff\x51\x03" + struct.pack(">I", TEMPO_US_PER_QUARTER)[1:])])]
    melodic_channels = [c for c in range(16) if c != 9]
    for gi, ((program, is_drum), ns) in enumerate(sorted(groups.items())):
        ch = 9 if is_drum else melodic_channels[gi % len(melodic_channels)]
        ev: List[Tuple[int, int, bytes]] = []
        if not is_drum:
            ev.append((0, 0, bytes([0xC0 | ch, program & 0x7F])))
        for n in ns:
            on = max(0, int(round(n.onset * TICKS_PER_SECOND)))
            dur = n.offset - n.onset if (math.isfinite(n.offset) and n.offset > n.onset) else default_duration
            off = max(on + 1, int(round((n.onset + dur) * TICKS_PER_SECOND)))
            ev.append((on, 2, bytes([0x90 | ch, n.pitch & 0x7F, velocity & 0x7F])))
            ev.append((off, 1, bytes([0x80 | ch, n.pitch & 0x7F, 0])))       # offs before ons at the same tick
        tracks.append(_track(ev))
    header = b"MThd" + struct.pack(">IHHH", 6, 1, len(tracks), TICKS_PER_QUARTER)
    return header + b"".join(tracks)


def write_midi(path, notes: Sequence[Note], **kw) -> None:
    with open(path, "wb") as f:
        f.write(notes_to_midi_bytes(notes, **kw))


def midi_bytes_to_notes(data: bytes) -> List[Note]:
    """Minimal reader for files produced by :func:`notes_to_midi_bytes` (note on/off, program change, tempo)."""
    if data[:4] != b"MThd":
        raise ValueError("not a MIDI file")
    _, fmt, ntracks, division = struct.unpack(">IHHH", data[4:14])
    pos, tempo, out = 14, TEMPO_US_PER_QUARTER, []
    for _ in range(ntracks):
        if data[pos:pos + 4] != b"MTrk":
            raise ValueError("bad track chunk")
        (length,) = struct.unpack(">I", data[pos + 4:pos + 8])
        p, end, tick = pos + 8, pos + 8 + length, 0
        program: Dict[int, int] = {}
        active: Dict[Tuple[int, int], float] = {}
        while p < end:
            delta = 0
            while True:
                b = data[p]
                p += 1
                delta = (delta << 7) | (b & 0x7F)
                if not b & 0x80:
                    break
            tick += delta
            status = data[p]
            if status == 0xFF:
                kind, ln = data[p + 1], data[p + 2]
                if kind == 0x51:
                    tempo = int.from_bytes(data[p + 3:p + 6], "big")
                p += 3 + ln
                continue
            ch, hi = status & 0x0F, status & 0xF0
            t = tick * tempo / (division * 1e6)
            if hi == 0xC0:
                program[ch] = data[p + 1]
                p += 2
            elif hi in (0x90, 0x80):
                pitch, vel = data[p + 1], data[p + 2]
                p += 3
                if hi == 0x90 and vel > 0:
                    active[(ch, pitch)] = t
                elif (ch, pitch) in active:
                    on = active.pop((ch, pitch))
                    out.append(Note(on, pitch, 128 if ch == 9 else program.get(ch, 0), ch == 9, t))
            else:
                raise ValueError(f"unsupported MIDI status 0x{status:02x}")
        pos = end
    out.sort(key=lambda n: (n.onset, n.is_drum, n.program, n.pitch))
    return out
