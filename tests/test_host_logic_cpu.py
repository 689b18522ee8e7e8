"""Host-side logic of the drop-in driver that needs no GPU (nerf/train_utils.py): the noise tensors are drawn in the
reference's order — per ray chunk rand[N,Nc], randn[N,Nc], rand[N,Nf], randn[N,Nc+Nf] (train_utils.py:75,
volume_rendering_utils.py:44, nerf_helpers.py:363), chunks in order — so that a seeded run consumes torch's generator exactly
as the unmodified reference does.  Checked against the draws recorded from the reference itself (tests/golden, written by
oracle/make_golden.py under torch.manual_seed(1234 + case index))."""
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASE_INDEX = {"stoch_train_64c64f": 2, "stoch_stress_chunks": 3}  # position in oracle/make_golden.py CASES -> seed 1234 + index


@pytest.mark.parametrize("case", sorted(CASE_INDEX))
def test_noise_draw_order_matches_the_reference(built_lib, case):
    from nerf import train_utils
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    n = int(g["H"]) * int(g["W"])
    opts = dict(num_coarse=int(g["num_coarse"]), num_fine=int(g["num_fine"]), perturb=bool(g["perturb"]),
                noise_std=float(g["noise_std"]), chunksize=int(g["chunksize"]))
    torch.manual_seed(1234 + CASE_INDEX[case])
    chunks = [train_utils._draw_noise(min(opts["chunksize"], n - st), opts, torch.device("cpu"), True)
              for st in range(0, n, opts["chunksize"])]
    got = train_utils._cat_noise(chunks)
    for key, name in (("t_rand", "noise_t_rand"), ("n_c", "noise_n_c"), ("u", "noise_u"), ("n_f", "noise_n_f")):
        assert torch.equal(got[key], torch.from_numpy(g[name])), key


def test_noise_is_skipped_when_the_reference_draws_none(built_lib):
    from nerf import train_utils
    opts = dict(num_coarse=8, num_fine=4, perturb=False, noise_std=0.0, chunksize=16)
    torch.manual_seed(3)
    before = torch.random.get_rng_state()
    out = train_utils._draw_noise(5, opts, torch.device("cpu"), True)
    assert all(v is None for v in out.values()) and torch.equal(before, torch.random.get_rng_state())
    out = train_utils._draw_noise(5, dict(opts, perturb=True), torch.device("cpu"), False)   # no fine network: no u / n_f draws
    assert out["t_rand"].shape == (5, 8) and out["u"] is None and out["n_f"] is None and out["n_c"] is None


def test_mode_options_and_unsupported_configs(built_lib):
    import nerf
    from nerf import train_utils
    blk = dict(num_coarse=64, num_fine=128, perturb=True, lindisp=False, radiance_field_noise_std=0.1, white_background=False, chunksize=2048)
    cfg = nerf.CfgNode(dict(nerf=dict(use_viewdirs=True, train=blk, validation=dict(blk, perturb=False, radiance_field_noise_std=0.0)),
                            dataset=dict(no_ndc=True, near=0.2, far=0.8)))
    o = train_utils._mode_opts(cfg, "train")
    assert (o["num_coarse"], o["num_fine"], o["perturb"], o["noise_std"], o["chunksize"]) == (64, 128, True, 0.1, 2048)
    assert train_utils._mode_opts(cfg, "validation")["perturb"] is False
    m = nerf.models.ConditionalBlendshapePaperNeRFModel(num_encoding_fn_xyz=10, num_encoding_fn_dir=4, include_input_xyz=True,
                                                        include_input_dir=False)
    rays = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="CUDA"):   # no CPU fallback: the product path refuses CPU tensors loudly
        nerf.run_one_iter_of_nerf(2, 2, 1.0, m, m, rays, rays, cfg, mode="train", expressions=torch.zeros(76), latent_code=torch.zeros(32))
    ndc = nerf.CfgNode(dict(nerf=dict(use_viewdirs=True, train=blk), dataset=dict(no_ndc=False, near=0.2, far=0.8)))
    with pytest.raises(NotImplementedError):
        nerf.run_one_iter_of_nerf(2, 2, 1.0, m, m, rays, rays, ndc, mode="train", expressions=torch.zeros(76), latent_code=torch.zeros(32))
