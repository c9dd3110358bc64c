"""the RDT -> controller hand-off kernel (vt_chunk_handoff) against the reference's own tensor-op sequence
(scripts/franka_model_eef.py:199-222,312 and scripts/franka_inference_eef.py:546,552-554), restated below with plain torch CPU ops;
bit-exact.  (The reference functions themselves cannot be imported: they sit in a module that needs RDT's un-vendored configs,
SigLIP and T5.)"""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vla_touch_b200 import native as nv  # noqa: E402
from vla_touch_b200 import rdt_handoff as hf  # noqa: E402


def reference_sequence(action, indices, T):
    joints = action[:, :, list(indices)]                                                       # franka_model_eef.py:211-212
    joints = joints * torch.tensor([[[1, 1, 1, 1, 1, 1, 1, 1, 1, 255]]], dtype=joints.dtype)   # :216-219
    vla_tensor = joints.to(torch.float32)                                                      # :312
    raw = vla_tensor.clone()                                                                   # what inference_fn copies to the host
    vla_tensor[:, :, -1] /= 255                                                                # franka_inference_eef.py:546
    return raw, vla_tensor[:, :T, :].contiguous()                                              # :553


def test_handoff_needs_a_cuda_tensor():
    with pytest.raises(nv.NativeError):
        hf.handoff_action_chunk(torch.zeros(1, 64, 128, dtype=torch.bfloat16))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("B,N,T", [(1, 64, 32), (8, 64, 64), (3, 17, 1)])
def test_rdt_handoff_is_bit_identical_to_the_reference_sequence(dtype, B, N, T):
    g = torch.Generator().manual_seed(B * 100 + N)
    action = (torch.randn(B, N, 128, generator=g) * torch.logspace(-3, 3, 128)).to(dtype)      # wide dynamic range: every rounding shows
    want_raw, want_chunk = reference_sequence(action, hf.RDT_EEF_INDICES, T)
    raw, chunk = hf.handoff_action_chunk(action.to("cuda:0"), T)
    assert raw.dtype == torch.float32 and chunk.shape == (B, T, 10)
    assert torch.equal(raw.cpu(), want_raw)
    assert torch.equal(chunk.cpu(), want_chunk)
    none, chunk2 = hf.handoff_action_chunk(action.to("cuda:0"), T, want_raw=False)
    assert none is None and torch.equal(chunk2, chunk)


@pytest.mark.gpu
def test_rdt_handoff_rejects_bad_requests():
    x = torch.zeros(1, 64, 128, dtype=torch.bfloat16, device="cuda:0")
    with pytest.raises(ValueError):
        hf.handoff_action_chunk(x, 65)
    with pytest.raises(IndexError):
        hf.handoff_action_chunk(x, 8, indices=(0, 1, 128))
    with pytest.raises(TypeError):
        hf.handoff_action_chunk(x.half(), 8)
    with pytest.raises(ValueError):
        hf.handoff_action_chunk(x[0], 8)


@pytest.mark.gpu
def test_handoff_feeds_predict_on_the_same_stream():
    """policy output (bf16, device) -> hand-off -> predict, no host copy in between."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import vt_testutil as U
    c = U.predict_case("predict_cfg1")
    ctl = U.make_controller(c, "cuda:0")
    T = c["T"]
    g = torch.Generator().manual_seed(2)
    action = torch.randn(1, 64, 128, generator=g).to(torch.bfloat16)
    _, want_chunk = reference_sequence(action, hf.RDT_EEF_INDICES, T)
    _, chunk = hf.handoff_action_chunk(action.to("cuda:0"), T)
    out = ctl.predict(c["state"][:1], chunk, c["img1"][:1], c["img2"][:1], c["forces"][:1])
    assert chunk.is_cuda and torch.equal(chunk.cpu(), want_chunk)
    assert out.shape == (1, T, c["A"]) and torch.isfinite(out).all()
