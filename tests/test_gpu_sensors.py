"""GPU parity of the Retina transform (bit-exact) and the odor-intensity sensor vs their numpy oracles."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_retina_bit_exact():
    import torch
    from flygym_b200.retina import Retina
    from oracle.retina_oracle import retina_oracle
    ret = Retina()
    rng = np.random.default_rng(0)
    n = 3
    img = rng.integers(0, 256, (n, 2, ret.H, ret.W, 3), dtype=np.uint8)
    img[1] = 255            # saturated frame: largest possible integer sums
    img[2, :, ::2] = 0      # structured frame
    out = ret(torch.from_numpy(img).cuda()).cpu().numpy()
    ref = retina_oracle(img, ret.id_map, ret.pale)
    assert out.shape == (n, 2, 721, 2)
    assert np.array_equal(out, ref), float(np.abs(out - ref).max())
    # saturated image -> every ommatidium reads exactly 1.0 in its own channel (up to one float32 rounding) and 0 elsewhere
    own = np.where(ret.pale[None, :], out[1, :, :, 1], out[1, :, :, 0])
    other = np.where(ret.pale[None, :], out[1, :, :, 0], out[1, :, :, 1])
    assert np.allclose(own, 1.0, atol=2e-7) and np.all(other == 0)
    # host-buffer variant gives the same bits
    out_h = np.empty_like(ref)
    ret.forward_host(img, out_h)
    assert np.array_equal(out_h, ref)
    # empty / ragged input handling
    with pytest.raises(ValueError):
        ret(torch.zeros((1, 2, ret.H, ret.W - 1, 3), dtype=torch.uint8, device="cuda"))


def test_retina_full_size_properties():
    """BASELINE config 4 size (1024 flies): linearity of the integer sums and mirror symmetry of the two eye maps."""
    import torch
    from flygym_b200.retina import Retina
    ret = Retina()
    n = 1024
    g = torch.Generator(device="cuda").manual_seed(0)
    img = torch.randint(0, 128, (n, 2, ret.H, ret.W, 3), dtype=torch.uint8, device="cuda", generator=g)
    a = ret(img)
    b = ret(img * 2)                      # values < 128 so doubling does not overflow uint8
    assert torch.allclose(b, 2 * a, rtol=0, atol=3e-7)
    mirrored = torch.flip(img[:, 0], dims=[2]).contiguous()      # eye 1's map is eye 0's mirror image
    both = torch.stack([img[:, 0], mirrored], dim=1).contiguous()
    c = ret(both)
    assert torch.equal(c[:, 0], c[:, 1])


def test_odor_sensor():
    import torch
    from flygym_b200 import B200Simulation
    from flygym_b200.retina import OdorSensor, ODOR_SENSORS
    from oracle.retina_oracle import odor_oracle
    sim = B200Simulation(None, n_worlds=5)
    sim.step(3)
    src = np.array([[3.0, 1.0, 0.5], [-2.0, 4.0, 1.0]], dtype=np.float32)
    peak = np.array([[1.0, 0.0], [0.3, 2.0]], dtype=np.float32)
    odor = OdorSensor(sim, src, peak)
    got = odor().cpu().numpy()
    segs = sim.model.names["segments"]
    ref = odor_oracle(sim.seg_xpos.cpu().numpy().astype(np.float64), sim.seg_xquat.cpu().numpy().astype(np.float64),
                      [segs.index(s) for s, _ in ODOR_SENSORS], [p for _, p in ODOR_SENSORS], src.astype(np.float64), peak.astype(np.float64))
    assert got.shape == (5, 2, 4)
    assert np.allclose(got, ref, rtol=1e-5, atol=0)


def test_eye_cameras_bit_exact_and_fused_path():
    """Eye-camera image formation: raw buffers bit-exact vs the float32 numpy restatement, and the fused
    render+Retina kernel identical to Retina(render)."""
    import torch
    from flygym_b200 import B200Simulation
    from flygym_b200.retina import EyeCameras
    from oracle.retina_oracle import eye_render_oracle, retina_oracle
    n = 6
    sim = B200Simulation(None, n_worlds=n)
    q = torch.tensor([1.0, 0.05, -0.1, 0.2]); q = q / q.norm()
    sim.qpos[1, 3:7] = q.cuda()                       # tilt one fly so that the horizon is not axis-aligned
    sim.qpos[2, 0:3] = torch.tensor([3.3, -1.7, 2.0]).cuda()
    # tumbling flies with bent legs: capsules at every angle to the image rows, next to / behind the cameras, vanishing points of
    # their axes inside the image (the silhouette-strip culling of the body raster takes all of its branches)
    g = torch.Generator().manual_seed(11)
    for i, quat in zip((3, 4, 5), ([0.8, 0.3, -0.4, 0.33], [0.1, 0.9, 0.2, -0.3], [0.5, -0.5, 0.6, 0.4])):
        qq = torch.tensor(quat); sim.qpos[i, 3:7] = (qq / qq.norm()).cuda()
        sim.qpos[i, 2] = 2.0
        sim.qpos[i, 7:] += (0.5 * torch.randn(sim.qpos.shape[1] - 7, generator=g)).cuda()
    sim.step(2)
    xp, xq = sim.seg_xpos.cpu().numpy(), sim.seg_xquat.cpu().numpy()
    plain = EyeCameras(sim, body=False)                  # ground and sky only (round-1 behaviour)
    ref0 = eye_render_oracle(xp, xq, plain.params, plain.ret.H, plain.ret.W)
    assert np.array_equal(plain.render().cpu().numpy(), ref0)
    cams = EyeCameras(sim)                               # + the fly's own body: the 55 segments outside the v1 hidden list, as capsules
    assert cams.body is not None and len(cams.body["seg"]) == 55
    img = cams.render()
    ref = eye_render_oracle(xp, xq, cams.params, cams.ret.H, cams.ret.W, body=cams.body)
    got = img.cpu().numpy()
    assert got.shape == (n, 2, 512, 450, 3)
    assert np.array_equal(got, ref), int((got != ref).sum())
    assert len(np.unique(got[..., 1])) >= 4              # both checker greys, the sky and the body are visible
    body_px = (got[..., 1] == cams.body["colour"][0]) & (got[..., 2] == cams.body["colour"][1])
    assert 0.03 < body_px.mean() < 0.5                   # legs / abdomen / wings cover part of the view, not all of it
    assert (got != ref0).any()
    fused = cams.retina()
    two_stage = cams.ret(img)
    assert torch.equal(fused, two_stage)
    assert np.array_equal(fused.cpu().numpy(), retina_oracle(ref, cams.ret.id_map, cams.ret.pale))
