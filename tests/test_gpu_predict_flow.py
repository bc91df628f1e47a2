"""The COMPOSITION of the predict path, bit for bit against the reference's own `predict.py __main__`.

tests/golden/predict_flow.npz holds what /root/reference/predict.py printed into its csv when run end to end (runpy,
oracle/gen_predict_flow.py) on the seeded clips and "ball detector" checkpoints of oracle/synth_clip.py, for every
eval_mode. Here the same clip and the same state_dicts go through this repo's `predict.run_video`: GPU median over the clip
-> Pillow-exact resize / stack -> TrackNet (tcgen05 kernels) -> [temporal ensemble] -> decode -> generate_inpaint_mask ->
InpaintNet + blend + threshold [-> temporal ensemble] -> integer coordinates. Frames, X, Y, Visibility and the inpaint
mask must be identical.
"""
import os

import numpy as np
import pytest
import torch

from oracle import synth_clip as S

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "predict_flow.npz")


def _models(seq_len, bg_mode):
    from utils.general import get_model
    tracknet = get_model('TrackNet', seq_len, bg_mode)
    tracknet.load_state_dict(S.detector_tracknet_state(seq_len, bg_mode, seed=1))
    inpaintnet = get_model('InpaintNet')
    inpaintnet.load_state_dict(S.inpaintnet_state(seed=2))
    return tracknet.cuda().eval(), inpaintnet.cuda().eval()


@pytest.mark.parametrize("case", S.CASES, ids=[c[0] for c in S.CASES])
def test_run_video_equals_reference_main(case):
    import predict as P
    name, t, (hs, ws), seq_len, bg_mode, l_inp, bs, gap = case
    g = np.load(GOLD)
    clip = torch.from_numpy(S.make_clip(t, hs, ws, seed=len(name) + t, gap=gap))
    tracknet, inpaintnet = _models(seq_len, bg_mode)
    for mode in S.EVAL_MODES:
        p1, p2 = P.run_video(clip, tracknet, inpaintnet, seq_len=seq_len, bg_mode=bg_mode, batch_size=bs, eval_mode=mode,
                             inpaintnet_seq_len=l_inp)
        for k in ("Frame", "X", "Y", "Visibility", "Inpaint_Mask"):
            assert list(g[f"{name}/{mode}/tracknet/{k}"]) == [int(v) for v in p1[k]], (name, mode, "tracknet", k)
        for k in ("Frame", "X", "Y", "Visibility"):
            assert list(g[f"{name}/{mode}/inpaint/{k}"]) == [int(v) for v in p2[k]], (name, mode, "inpaint", k)
    # the fixture is not vacuous: the ball is found in most frames, lost in the gap, and the gap is what gets inpainted
    vis = g[f"{name}/weight/tracknet/Visibility"]
    assert vis.sum() == t - len(gap) and all(vis[f] == 0 for f in gap)
    assert list(np.nonzero(g[f"{name}/weight/tracknet/Inpaint_Mask"])[0]) == list(gap)


@pytest.mark.parametrize("t", [1, 2, 5, 6, 31])
def test_median_kernel_equals_numpy(t):
    """tnb_median_u8 vs np.median(frame_arr, 0) (dataset.py:103): float64 result incl. the x.5 values of even clip
    lengths, and its .astype('uint8') (dataset.py:105); ragged sizes exercise the byte tail."""
    import tracknetv3_b200 as T
    rs = np.random.RandomState(t)
    for hs, ws in ((6, 10), (7, 9), (5, 3)):
        frames = rs.randint(0, 256, size=(t, hs, ws, 3)).astype(np.uint8)
        frames[:, 0, 0] = 255
        frames[:, 1, 1] = 0
        fp = T.FramePreprocessor(hs, ws, 288, 512)
        want = np.median(frames, 0)
        got_f = fp.median(torch.from_numpy(frames).cuda(), as_float=True).cpu().numpy()
        got_u = fp.median(torch.from_numpy(frames).cuda()).cpu().numpy()
        assert got_f.dtype == np.float64 and np.array_equal(got_f, want)
        assert np.array_equal(got_u, want.astype('uint8'))


def test_short_clip_gives_the_reference_empty_result():
    """Fewer frames than seq_len: the reference's dataset has no sample without padding (dataset.py:338-353), its loops do
    not run and the dictionaries stay empty; with eval_mode 'nonoverlap' the single padded window covers every frame."""
    import predict as P
    tracknet, inpaintnet = _models(8, 'concat')
    clip = torch.from_numpy(S.make_clip(5, 288, 512, seed=3, gap=()))
    p1, p2 = P.run_video(clip, tracknet, inpaintnet, eval_mode='weight', inpaintnet_seq_len=4)
    assert p1['Frame'] == [] and p2['Frame'] == []
    p1, p2 = P.run_video(clip, tracknet, inpaintnet, eval_mode='nonoverlap', inpaintnet_seq_len=4)
    assert p1['Frame'] == [0, 1, 2, 3, 4] and p2['Frame'] == [0, 1, 2, 3, 4] and sum(p1['Visibility']) == 5


def test_inpaintnet_rectify_equals_separate_steps():
    """tnb_inpaintnet_rectify (blend + COOR_TH in the kernel epilogue) == forward kernel followed by the reference's
    torch expressions (predict.py:257-261), bit for bit."""
    from utils.general import get_model, COOR_TH
    torch.manual_seed(4)
    net = get_model('InpaintNet').cuda().eval()
    c = torch.rand(9, 16, 2).cuda()
    c[torch.rand(9, 16) < 0.3] = 0.01
    m = (torch.rand(9, 16, 1) < 0.4).float().cuda()
    with torch.no_grad():
        out = net(c, m)
        want = out * m + c * (1 - m)
        want[(want[:, :, 0] < COOR_TH) & (want[:, :, 1] < COOR_TH)] = 0.
    assert torch.equal(net.rectify(c, m, COOR_TH), want)
