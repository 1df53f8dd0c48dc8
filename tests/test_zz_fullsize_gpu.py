"""GPU, last in collection order: end-to-end parity at the BENCH configuration (256x512, D = 192), where every tensor-core
route is live at once -- transposed-image backbone front, layer2/3/4 residual blocks and lastconv on tcgen05, fused volume,
channels-last aggregation, fused tail -- against the CPU oracle with the same seeded weights (one pair; the oracle needs a
second or two on the box's host cores).  Bar: the north star's 1e-3 px EPE."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import models as omodels       # noqa: E402
from oracle import seeded_init as si       # noqa: E402


def test_gwcnet_bench_configuration_epe():
    import __graft_entry__
    __graft_entry__.build()
    from openstereo_b200 import _lib, host_models
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = {"MAX_DISP": 192, "USE_CONCAT_VOLUME": True, "CONCAT_CHANNELS": 12, "DOWNSAMPLE": 4, "NUM_GROUPS": 40}
    oracle = omodels.GwcNet(192, True, 12, 4, 40).eval()
    sd = si.seeded_state_dict(oracle.state_dict(), seed=1, scale=si.GWCNET_SCALE)
    oracle.load_state_dict(sd)
    mine = host_models.GwcNet(cfg).eval()
    mine.load_state_dict(sd)
    mine.cuda()
    g = torch.Generator().manual_seed(0)
    x = {"left": torch.randn(1, 3, 256, 512, generator=g), "right": torch.randn(1, 3, 256, 512, generator=g)}
    with torch.no_grad():
        want = oracle(dict(x))["disp_pred"]
        before = _lib.launch_count()
        got = mine({k: v.cuda() for k, v in x.items()})["disp_pred"]
        launches = _lib.launch_count() - before
    assert got.shape == want.shape == (1, 256, 512)
    e = (got.cpu() - want).abs().mean().item()
    print("GwcNet 256x512 (bench configuration) EPE vs oracle: %.3e px, %d launches of this library" % (e, launches))
    assert launches >= 80                      # the tensor-core backbone routes (49 launches) are taken, not cuDNN's
    assert want.std() > 10 and e <= 1e-3
