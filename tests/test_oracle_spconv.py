"""Self-validation of the sparse-conv oracle (parity unpinned at the spconv boundary):
every layer type the Det3D encoders use must equal dense conv3d at the active output set."""
import numpy as np
import pytest
import torch

from oracle import spconv as osp


def _random_sites(rng, n, spatial, batch):
    d, h, w = spatial
    cells = rng.choice(batch * d * h * w, n, replace=False)
    b, r = cells // (d * h * w), cells % (d * h * w)
    return np.stack([b, r // (h * w), (r // w) % h, r % w], 1).astype(np.int32)


@pytest.mark.parametrize("k,s,p,subm", [(3, 1, 1, True), (3, 2, 1, False), (3, 2, [0, 1, 1], False),
                                        ((3, 1, 1), (2, 1, 1), 0, False), (1, 1, 0, True)])
def test_layer_equals_dense_conv3d(k, s, p, subm):
    rng = np.random.default_rng(7)
    spatial, batch = (9, 14, 11), 2
    coors = _random_sites(rng, 220, spatial, batch)
    feat = rng.standard_normal((220, 6)).astype(np.float32)
    kk = osp._triple(k)
    w = (rng.standard_normal((*kk, 6, 8)) * 0.2).astype(np.float32)
    err = osp.check_against_dense(feat, coors, spatial, batch, w, k, s, p, subm, bias=rng.standard_normal(8).astype(np.float32))
    assert err < 1e-5


def test_conv_outputs_sorted_and_unique():
    rng = np.random.default_rng(3)
    spatial = (11, 20, 16)
    coors = _random_sites(rng, 300, spatial, 3)
    oc, osp_ = osp.conv_outputs(coors, spatial, 3, 2, 1)
    lin = osp.linear_index(oc, osp_)
    assert (np.diff(lin) > 0).all()
    assert osp_ == (6, 10, 8)


def test_subm_centre_is_identity_and_symmetric():
    rng = np.random.default_rng(4)
    spatial = (8, 8, 8)
    coors = _random_sites(rng, 100, spatial, 1)
    nbr = osp.subm_neighbours(coors, spatial, 3)
    assert (nbr[13] == np.arange(100)).all()
    # (k, i -> o) exists iff (26-k, o -> i) exists
    for k in range(27):
        o = np.nonzero(nbr[k] >= 0)[0]
        assert (nbr[26 - k][nbr[k][o]] == o).all()


def test_middle_encoder_shapes():
    import sys
    from det3d_b200.models.backbones.scn import SpMiddleFHD, SpMiddleResNetFHD
    from det3d_b200.utils.synthetic import randomize_bn_
    torch.manual_seed(0)
    rng = np.random.default_rng(0)
    input_shape = [64, 80, 40]                      # x, y, z
    spatial = (17, 80, 64)
    coors = _random_sites(rng, 400, (40, 80, 64), 2)
    for cls, cin, cout in ((SpMiddleFHD, 4, 64), (SpMiddleResNetFHD, 5, 128)):
        m = randomize_bn_(cls(num_input_features=cin).eval())
        feats = rng.standard_normal((400, cin)).astype(np.float32)
        out = osp.middle_encoder_forward(m.state_dict(), feats, coors, 2, input_shape, arch=cls.__name__)
        assert out.shape == (2, cout * 2, 10, 8)
        assert torch.isfinite(out).all() and out.abs().sum() > 0
