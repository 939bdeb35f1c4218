"""CPU: checkpoint compatibility of the model mirrors with the reference (run.py:786-793, lib/utils.py:232-300 load
with strict=True).  tests/golden/r2_extras.npz holds the sorted "key:shape" schema of the reference's own fine and coarse
models' state_dict()."""
import torch

from voxurf_b200 import synthetic as S
from tests.helpers import load_golden, mask_cache_state


def _sd_from_schema(schema):
    sd = {}
    for item in schema:
        k, shp = str(item).split(':')
        shape = tuple(int(v) for v in shp.strip('()').split(',') if v.strip())
        sd[k] = torch.zeros(shape, dtype=torch.bool) if k == 'nonempty_mask' else torch.full(shape, 0.25)
    return sd


def test_fine_state_dict_schema_matches_reference():
    from voxurf_b200 import voxurf_fine as VF
    g = load_golden('r2_extras.npz')
    sc = S.make_fine_scene(20, 6, 32, seed=3, mask_G=12)
    cfg = {k: v for k, v in S.FINE_CFG.items() if k != 'stepsize'}
    m = VF.Voxurf(xyz_min=[-1., -1., -1.], xyz_max=[1., 1., 1.], num_voxels=20 ** 3, num_voxels_base=20 ** 3, rgbnet_dim=6,
                  rgbnet_width=32, smooth_ksize=5, smooth_sigma=0.8, mask_cache_state=mask_cache_state(sc), **cfg)
    sd = _sd_from_schema(g['schema_fine'])
    sd['nonempty_mask'][0, 0, 3:9, 2:5, 1:4] = True
    sd['s_val'][:] = 0.125
    m.load_state_dict(sd, strict=True)       # raises on any missing / unexpected key or shape mismatch
    assert sorted(m.state_dict().keys()) == sorted(sd.keys())
    # derived host-side values follow the loaded tensors
    assert m._n_nonempty == 6 * 3 * 3 and m._s_val_host == 0.125
    assert not any(p.requires_grad for n, p in m.named_parameters() if 'conv' in n)


def test_coarse_state_dict_schema_matches_reference():
    from voxurf_b200 import voxurf_coarse as VC
    g = load_golden('r2_extras.npz')
    sc = S.make_coarse_scene(16, 12, 32, seed=4, mask_G=12)
    cfg = {k: v for k, v in S.COARSE_CFG.items() if k != 'stepsize'}
    cfg['rgbnet_dim'], cfg['rgbnet_width'] = 12, 32
    m = VC.Voxurf(xyz_min=[-1., -1., -1.], xyz_max=[1., 1., 1.], num_voxels=16 ** 3, num_voxels_base=16 ** 3,
                  rgbnet_direct=True, mask_cache_state=mask_cache_state(sc), **cfg)
    sd = _sd_from_schema(g['schema_coarse'])
    m.load_state_dict(sd, strict=True)
    assert sorted(m.state_dict().keys()) == sorted(sd.keys())
