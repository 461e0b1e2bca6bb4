"""Host-side logic of the packed-row encoder and of the captured train step (no GPU needed):
row/gather index maps, the device-int cache, and the RAdam / Noam schedule `FusedRAdam.advance`
hands to the kernels (against the oracle's restatement of Radam.py:57-76 / Noam_Scheduler.py:17-29)."""
import math

import numpy as np
import torch


def test_token_rows_pack_unpack_roundtrip_cpu():
    from glow_tts_b200 import rows
    lens, t_max = [7, 1, 12, 3], 12
    tr = rows.TokenRows(lens, t_max, torch.device("cpu"))
    assert tr.rows_pad % 128 == 0 and tr.rows_pad >= sum(lens) + 2 * (len(lens) + 1)
    row_utt = tr.rm.row_utt.numpy()
    assert (row_utt[:2] == -1).all() and (row_utt[-2:] == -1).all()           # leading / trailing guard rows
    for b, n in enumerate(lens):
        off = int(tr.rm.utt_off[b])
        assert (row_utt[off:off + n] == b).all()
        assert (row_utt[off - 2:off] == -1).all() and (row_utt[off + n:off + n + 2] == -1).all()
    x = torch.randn(len(lens), t_max, 5)
    packed = tr.pack(x)
    assert packed.shape == (tr.rows_pad, 5)
    assert float((packed * (1 - tr.valid)).abs().max()) == 0.0                # guard rows are zero
    back = tr.unpack(packed)
    mask = tr.tmask.unsqueeze(2)
    assert torch.equal(back, x * mask)                                        # exact on real tokens, zero beyond
    # a row holds the token (b, t) its map says
    for r in np.flatnonzero(row_utt >= 0)[::5]:
        b, t = int(row_utt[r]), int(tr.rm.row_t[r])
        assert torch.equal(packed[r], x[b, t])


def test_device_ints_are_cached():
    from glow_tts_b200 import _lib
    a = _lib.device_ints([3, 1, 2], torch.int32, "cpu")
    b = _lib.device_ints((3, 1, 2), torch.int32, torch.device("cpu"))
    assert a is b and a.dtype == torch.int32 and a.tolist() == [3, 1, 2]
    assert _lib.device_ints([3, 1, 2], torch.int64, "cpu") is not a


def test_fused_radam_schedule_matches_oracle():
    from glow_tts_b200.flat import FlatBuffer
    from glow_tts_b200.train import FusedRAdam
    from oracle.glow_oracle import RAdamOracle
    p = torch.nn.Parameter(torch.zeros(8))
    opt = FusedRAdam(FlatBuffer([p]), lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=1e-6, base=4000, max_norm=5.0)
    ref = RAdamOracle([torch.zeros(8)])
    for step in range(1, 12):
        lr, b1, b2, eps, wd, step_size, rect, max_norm, grad_scale = opt.advance(grad_scale=0.5)
        # the oracle's scalars for the same step
        ref.t += 1
        b2t = 0.999 ** ref.t
        n_max = 2 / (1 - 0.999) - 1
        n_sma = n_max - 2 * ref.t * b2t / (1 - b2t)
        if n_sma >= 5:
            want = math.sqrt((1 - b2t) * (n_sma - 4) / (n_max - 4) * (n_sma - 2) / n_sma * n_max / (n_max - 2)) / (1 - 0.9 ** ref.t)
        else:
            want = 1.0 / (1 - 0.9 ** ref.t)
        assert lr == ref.lr() and step_size == want and rect == float(n_sma >= 5)
        assert (b1, b2, eps, wd, max_norm, grad_scale) == (0.9, 0.999, 1e-6, 1e-6, 5.0, 0.5)
        ref.epoch += 1
    assert opt.steps == 11 and opt.epoch == 11


def test_device_row_map_matches_host_row_map_semantics():
    """flow.DeviceRowMap (fixed geometry, validity computed from device-resident lengths -- the sync-free inference
    path) against the invariants of the host-built flow.RowMap: guard rows, per-utterance runs, frame indices."""
    from glow_tts_b200 import flow
    batch, sq_max = 5, 37
    rm = flow.DeviceRowMap(batch, sq_max, "cpu")
    assert rm.rows_pad % flow.ROW_TILE == 0 and rm.rows_pad >= flow.GUARD + batch * (sq_max + flow.GUARD)
    for lens in ([37, 0, 12, 1, 30], [5, 5, 5, 5, 5], [99, 37, 36, 0, 0]):          # 99 is clamped to sq_max
        rm.update(torch.tensor(lens, dtype=torch.int64))
        want = [min(n, sq_max) for n in lens]
        assert rm.utt_len.tolist() == want
        row_utt, row_t = rm.row_utt.numpy(), rm.row_t.numpy()
        assert int((row_utt >= 0).sum()) == sum(want)
        assert (row_utt[:flow.GUARD] == -1).all()
        for b, n in enumerate(want):
            off = int(rm.utt_off[b])
            assert (row_utt[off:off + n] == b).all()
            assert (row_t[off:off + n] == np.arange(n)).all()
            assert (row_utt[off + n:off + sq_max + flow.GUARD] == -1).all()          # rest of the slot + its guard rows
            assert (row_utt[off - flow.GUARD:off] == -1).all()
        # same runs as the host map of the same lengths, only placed on the fixed stride
        host = flow.RowMap(want, "cpu")
        assert host.utt_len.tolist() == want
        for b, n in enumerate(want):
            ho, do = int(host.utt_off[b]), int(rm.utt_off[b])
            assert (host.row_t.numpy()[ho:ho + n] == row_t[do:do + n]).all()


def test_step_geometry_host_arrays_match_row_map():
    """geometry.fill_row_arrays (static, bucketed buffers) writes the same maps flow.RowMap builds per batch."""
    from glow_tts_b200 import flow, geometry
    lens = [17, 3, 40, 1]
    rm = flow.RowMap(lens, torch.device("cpu"))
    rows_pad = 256
    utt, t, off, ln = (np.zeros(rows_pad, np.int32), np.zeros(rows_pad, np.int32), np.zeros(4, np.int32),
                       np.zeros(4, np.int32))
    geometry.fill_row_arrays(lens, rows_pad, utt, t, off, ln)
    n = rm.rows_pad
    assert np.array_equal(utt[:n], rm.row_utt.numpy()) and (utt[n:] == -1).all()
    assert np.array_equal(t[:n], rm.row_t.numpy())
    assert np.array_equal(off, rm.utt_off.numpy()) and np.array_equal(ln, rm.utt_len.numpy())
    assert geometry.rows_needed(lens) == 2 + sum(lens) + 2 * len(lens)
    try:
        geometry.fill_row_arrays([200, 200], 256, utt, t, off[:2], ln[:2])
        assert False, "overflow must be refused"
    except ValueError:
        pass


def test_bucket_of_rounds_rows_up():
    from glow_tts_b200.geometry import StepGeometry, DEC_ROW_BUCKET, ENC_ROW_BUCKET
    tl, ml = [50, 20, 33], [400, 150, 260]
    key = StepGeometry.bucket_of(tl, ml, 202, 1000)
    b, rd, re, tt, tm = key
    assert b == 3 and tt == 202 and tm == 1000
    assert rd % DEC_ROW_BUCKET == 0 and rd >= 2 + sum(n // 2 + 2 for n in ml) > rd - DEC_ROW_BUCKET
    assert re % ENC_ROW_BUCKET == 0 and re >= 2 + sum(n + 2 for n in tl) > re - ENC_ROW_BUCKET
    # a slightly different batch falls into the same bucket, a much longer one does not
    assert StepGeometry.bucket_of([49, 22, 31], [396, 158, 262], 202, 1000) == key
    assert StepGeometry.bucket_of(tl, [1000, 900, 800], 202, 1000) != key


def test_fused_radam_state_dict_uses_the_reference_layout():
    """FusedRAdam.state_dict() / load_state_dict() speak the reference checkpoint's optimizer format
    (Train.py:514-519: torch.optim layout of Radam.py, one state entry per parameter in model.parameters() order,
    plus the scheduler's last_epoch): a reference-format dict loads into the flat buffers at the right offsets and
    the round trip is exact."""
    from glow_tts_b200.flat import FlatBuffer
    from glow_tts_b200.train import FusedRAdam
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(3, 5)), torch.nn.Parameter(torch.randn(7)), torch.nn.Parameter(torch.randn(2, 2, 3))]
    flat = FlatBuffer(params)
    opt = FusedRAdam(flat)
    ref_state = {i: {"step": 41, "exp_avg": torch.randn_like(p), "exp_avg_sq": torch.rand_like(p)} for i, p in enumerate(params)}
    ref_sd = {"state": ref_state, "param_groups": [{"lr": 9.9e-4, "betas": (0.9, 0.999), "eps": 1e-6, "weight_decay": 1e-6,
                                                     "initial_lr": 1e-3, "params": [0, 1, 2]}]}
    sch_sd = {"base": 4000, "base_lrs": [1e-3], "last_epoch": 41, "_step_count": 42, "_last_lr": [9.9e-4]}
    opt.load_state_dict(ref_sd, sch_sd)
    assert opt.steps == 41 and opt.epoch == 41
    for i, (p, o) in enumerate(zip(flat.params, flat.offsets)):
        assert torch.equal(opt.exp_avg[o:o + p.numel()].view(p.shape), ref_state[i]["exp_avg"])
        assert torch.equal(opt.exp_avg_sq[o:o + p.numel()].view(p.shape), ref_state[i]["exp_avg_sq"])
    assert abs(opt.lr() - 1e-3 * 4000 ** 0.5 * (41 + 4000) ** -0.5) < 1e-12       # Noam position restored
    out = opt.state_dict()
    assert sorted(out["state"].keys()) == [0, 1, 2] and out["param_groups"][0]["params"] == [0, 1, 2]
    for i in range(3):
        assert out["state"][i]["step"] == 41
        assert torch.equal(out["state"][i]["exp_avg"], ref_state[i]["exp_avg"])
        assert torch.equal(out["state"][i]["exp_avg_sq"], ref_state[i]["exp_avg_sq"])
    assert opt.scheduler_state_dict()["last_epoch"] == 41
    # torch.optim.Adam accepts the same dict shape (the layout really is torch.optim's)
    torch.optim.Adam(params).load_state_dict({"state": {i: dict(v, step=torch.tensor(41.0)) for i, v in out["state"].items()},
                                              "param_groups": [dict(torch.optim.Adam(params).state_dict()["param_groups"][0])]})
    # the next step continues the schedule: step 42 is past the N_sma >= 5 threshold, so it is rectified
    hyper = opt.advance()
    assert opt.steps == 42 and hyper[6] == 1.0
