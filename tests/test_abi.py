"""CPU: the C-ABI shared library builds for sm_100a, loads, and exports every symbol include/egaze.h declares
(no compute calls without a GPU); the drop-in modules keep the reference's state layout."""
import ctypes
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    return ge.build()


def test_library_exports_header(built):
    from egaze import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 25
    h = ctypes.CDLL(built)
    for name in protos:
        assert hasattr(h, name), "missing export %s" % name
    assert h.egaze_version() >= 100


def test_error_convention(built):
    from egaze import _lib
    h = _lib.lib()
    # invalid argument -> negative rc + message; no exception crosses the ABI, nothing launched
    rc = h.egaze_bn_finalize(None, None, 1, 1, 0, 0, 1e-5, 0.1, None, None, None, None, None, None, None, None, None, None)
    assert rc < 0
    assert "bn_finalize" in _lib.last_error()


def test_no_cpu_fallback(built):
    """The product path must fail loudly without a CUDA device (never route through the oracle / PyTorch)."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from utils import make_layers, cfg
    from models.model_SP import model_SP
    from models.late_fusion import late_fusion
    from models.LSTMnet import lstmnet
    from floss import floss
    m = model_SP(make_layers(cfg['D'], 3), make_layers(cfg['D'], 20)).eval()
    with torch.no_grad():
        with pytest.raises(RuntimeError):
            m(torch.zeros(1, 3, 32, 32), torch.zeros(1, 20, 32, 32))
        with pytest.raises(RuntimeError):
            late_fusion().eval()(torch.zeros(1, 1, 32, 32), torch.zeros(1, 1, 32, 32))
        with pytest.raises(RuntimeError):
            lstmnet()(torch.zeros(1, 1, 512), None)
        with pytest.raises(RuntimeError):
            floss()(torch.rand(1, 1, 8, 8), torch.rand(1, 1, 8, 8))


def test_state_layout_matches_oracle_description(built):
    """215 keys for model_SP (t-first), 23 for late_fusion, lstm + lin keys for lstmnet."""
    from test_oracle_golden import sp_shapes, lf_shapes, lstm_shapes
    from utils import make_layers, cfg
    from models.model_SP import model_SP
    from models.late_fusion import late_fusion
    from models.LSTMnet import lstmnet
    m = model_SP(make_layers(cfg['D'], 3), make_layers(cfg['D'], 20))
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == list(sp_shapes().items())
    assert [n for n, _ in m.named_children()] == ['features_t', 'features_s', 'relu', 'fusion', 'pool3d', 'bn', 'decoder', 'final']
    assert [(k, tuple(v.shape)) for k, v in late_fusion().state_dict().items()] == list(lf_shapes().items())
    assert {k: tuple(v.shape) for k, v in lstmnet().state_dict().items()} == lstm_shapes()
    conv_idx = [i for i, mod in enumerate(m.features_s) if isinstance(mod, torch.nn.Conv2d)]
    assert conv_idx == [0, 3, 7, 10, 14, 17, 20, 24, 27, 30, 34, 37, 40]
    dec_idx = [i for i, mod in enumerate(m.decoder) if isinstance(mod, torch.nn.Conv2d)]
    assert dec_idx == [0, 2, 5, 7, 9, 12, 14, 16, 19, 21, 24, 26, 28]
