"""checkpoint adapter (Lightning-style key prefixes, tied LM head, shape checks) and the
dependency-free MIDI writer (round trip through the minimal reader; header / delta-time encoding known answers)."""
import math

import pytest
import torch

import yourmt3_b200 as ymt3
from yourmt3_b200 import checkpoint as CK
from yourmt3_b200 import midi as MD
from yourmt3_b200.event_codec import Note, onset_f1


def _tiny():
    cfg = ymt3.get_model_cfg("mt3_t5_small")
    cfg["encoder"]["t5"]["num_layers"] = cfg["decoder"]["t5"]["num_layers"] = 1
    cfg["event_length"] = 8
    return ymt3.YourMT3(model_cfg=cfg, precision="f32")


def test_lightning_checkpoint_round_trip():
    src, dst = ymt3.init_nondegenerate_(_tiny(), 3), _tiny()
    sd = src.state_dict()
    ckpt = {"state_dict": {("model." + k): v.to(torch.bfloat16) for k, v in sd.items() if k != "lm_head.lm_head.weight"},
            "epoch": 7}
    ckpt["state_dict"]["model.loss_weight"] = torch.ones(3)          # training-only entry: dropped silently
    missing, unexpected = CK.load_checkpoint(dst, ckpt)
    assert missing == [] and unexpected == []
    for k, v in dst.state_dict().items():
        assert v.dtype == sd[k].dtype
        assert torch.equal(v, sd[k].to(torch.bfloat16).to(v.dtype)), k
    assert dst.lm_head.lm_head.weight.data_ptr() == dst.embed_tokens.weight.data_ptr() or torch.equal(
        dst.lm_head.lm_head.weight, dst.embed_tokens.weight)


def test_checkpoint_mismatches_are_reported():
    m = _tiny()
    sd = {("module.model." + k): v for k, v in m.state_dict().items()}
    first = next(k for k in sd if k.endswith("SelfAttention.q.weight"))
    extra = dict(sd)
    extra["module.model.encoder.block.9.layer.0.SelfAttention.q.weight"] = torch.zeros(2, 2)
    with pytest.raises(KeyError):
        CK.load_checkpoint(m, extra)
    missing, unexpected = CK.load_checkpoint(m, extra, strict=False)
    assert missing == [] and unexpected == ["encoder.block.9.layer.0.SelfAttention.q.weight"]
    bad = dict(sd)
    bad[first] = torch.zeros(3, 3)
    with pytest.raises(ValueError):
        CK.load_checkpoint(m, bad)
    del sd[first]
    missing, _ = CK.load_checkpoint(m, sd, strict=False)
    assert missing == [first[len("module.model."):]]


def test_midi_known_answers():
    assert MD._vlq(0) == b"\x00" and MD._vlq(0x7F) == b"\x7f" and MD._vlq(0x80) == b"\x81\x00"
    assert MD._vlq(0x3FFF) == b"\xff\x7f" and MD._vlq(0x200000) == b"\x81\x80\x80\x00"      # SMF spec examples
    data = MD.notes_to_midi_bytes([Note(0.5, 60, 0, False, 1.0)])
    assert data[:14] == b"MThd\x00\x00\x00\x06\x00\x01\x00\x02\x01\xe0"                     # format 1, 2 tracks, 480 tpq
    assert b"\xff\x51\x03\x07\xa1\x20" in data                                              # 500000 us / quarter
    assert b"\x83\x60\x90\x3c\x64" in data                                                  # delta 480 ticks, note on C4


def test_midi_round_trip_and_metrics():
    g = torch.Generator().manual_seed(5)
    notes = []
    for i in range(200):
        onset = float(torch.rand(1, generator=g)) * 30.0
        drum = bool(torch.rand(1, generator=g) < 0.2)
        dur = float("nan") if (drum or i % 7 == 0) else 0.05 + float(torch.rand(1, generator=g))
        notes.append(Note(onset, int(torch.randint(21, 108, (1,), generator=g)), 128 if drum else int(i % 5) * 8, drum,
                          onset + dur))
    back = MD.midi_bytes_to_notes(MD.notes_to_midi_bytes(notes))
    # overlapping notes of the same pitch on one channel legitimately merge on the way back: compare by onsets
    p, r, f = onset_f1(notes, back, tolerance=0.002)
    assert r > 0.97 and p > 0.97 and f > 0.97
    for n in back:
        assert math.isfinite(n.offset) and n.offset > n.onset
    progs = {(n.program, n.is_drum) for n in back}
    assert progs == {(n.program, n.is_drum) for n in notes}


def test_load_checkpoint_file_with_pickled_hparams(tmp_path):
    """Lightning-style file: non-tensor objects next to the state dict -> clear error by default, loads with
    trusted=True; upstream-only buffers (rotary tables, num_batches_tracked) are ignored, not 'unexpected'."""
    import argparse
    import pytest
    import torch
    import yourmt3_b200 as ymt3
    from yourmt3_b200 import checkpoint as CK
    cfg = ymt3.get_model_cfg("mt3_t5_small")
    cfg["encoder"]["t5"]["num_layers"] = cfg["decoder"]["t5"]["num_layers"] = 1
    m = ymt3.YourMT3(model_cfg=cfg)
    sd = {"model." + k: v.clone() for k, v in m.state_dict().items()}
    sd["model.encoder.rotary.inv_freq"] = torch.ones(4)
    sd["model.pre_encoder.bn.num_batches_tracked"] = torch.tensor(3)
    p = tmp_path / "lightning.ckpt"
    torch.save({"state_dict": sd, "hyper_parameters": argparse.Namespace(lr=1e-3), "epoch": 3}, p)
    with pytest.raises(RuntimeError, match="trusted=True"):
        CK.load_checkpoint(m, str(p))
    missing, unexpected = CK.load_checkpoint(m, str(p), trusted=True)
    assert missing == [] and unexpected == []


def test_midi_bytes_against_hand_assembled_file():
    """Independent of the module's own reader: the exact bytes of a two-note file, assembled by hand from the
    Standard MIDI File 1.0 specification (header chunk, tempo meta event FF 51 03, program change Cn, note on 9n /
    note off 8n, variable-length delta times, end-of-track FF 2F 00).  120 bpm, 480 ticks per quarter = 960 ticks/s."""
    from yourmt3_b200 import midi as MD
    from yourmt3_b200.event_codec import Note
    notes = [Note(0.0, 60, 5, False, 0.5),          # program 5: ticks 0 .. 480
             Note(0.25, 64, 5, False, 1.0),         #            ticks 240 .. 960
             Note(0.5, 38, 128, True)]              # drum, default duration 0.1 s: ticks 480 .. 576
    got = MD.notes_to_midi_bytes(notes)
    tempo = bytes.fromhex("4d54726b" "0000000b" "00ff5103" "07a120" "00ff2f00")
    # tracks are ordered by (program, is_drum): the drum group (0, True) comes first, the program-5 group is the
    # second group and takes the second melodic channel (index 1)
    # drum track on channel 9: dt 480 (83 60) 99 26 64 | dt 96 (60) 89 26 00
    drum_body = bytes.fromhex("8360992664" "60892600" "00ff2f00")
    # melodic track on channel 1: dt 0 C1 05 | dt 0 91 3C 64 | dt 240 (81 70) 91 40 64 | dt 240 81 3C 00 | dt 480 (83 60) 81 40 00
    mel_body = bytes.fromhex("00c105" "00913c64" "8170914064" "8170813c00" "8360814000" "00ff2f00")
    chunk = lambda body: b"MTrk" + len(body).to_bytes(4, "big") + body
    want = bytes.fromhex("4d546864" "00000006" "0001" "0003" "01e0") + tempo + chunk(drum_body) + chunk(mel_body)
    assert got == want, (got.hex(), want.hex())
    # variable-length quantities at the boundaries the specification lists
    assert [MD._vlq(v).hex() for v in (0, 0x7F, 0x80, 0x2000, 0x3FFF, 0x4000, 0x0FFFFFFF)] == \
        ["00", "7f", "8100", "c000", "ff7f", "818000", "ffffff7f"]
