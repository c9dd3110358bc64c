"""Developer tool (GPU box): one small invocation of each hand-written tensor-core kernel, to be run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize.py > gpurun_out/sanitizer_memcheck.log
    compute-sanitizer --tool racecheck python tools/sanitize.py > gpurun_out/sanitizer_racecheck.log

gemm_tc_kernel (linear + GroupNorm epilogues, CTA pairs), mlp_fused_kernel, attn_row_kernel, unet_persist_kernel (whole sampler),
wgrad_tc_kernel, lstm_tc_kernel / lstm_bwd_tc_kernel, batch_gather_kernel, chunk_handoff_kernel -- through the public API at the smallest shapes that reach each kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402

import vt_testutil as U  # noqa: E402
from vla_touch_b200 import shapes as shp  # noqa: E402
from vla_touch_b200 import synthetic as syn  # noqa: E402


def main():
    dev = "cuda:0"
    c = U.predict_case("predict_cfg2_B3_dark_varstats")          # 2 ViT layers, 224 x 224, batch 3, T 64, 10 SDE steps
    ctl = U.make_controller(c, dev)
    ctl.noise_override = c["gold"]["noise"].to(dev)
    out = ctl.predict(c["state"].to(dev), c["vla"].to(dev), c["img1"], c["img2"], c["forces"].to(dev))
    torch.cuda.synchronize()
    print("predict (gemm_tc, mlp_fused, attn_row, unet_persist):", tuple(out.shape), float(out.abs().max()))
    si = ctl.diffusion_model
    cond = ctl.encode_observation(c["state"].to(dev), c["img1"], c["img2"], c["forces"].to(dev))
    loss, _ = si.get_loss({"obs_cond": cond, "expert_act": torch.zeros(3, c["T"], c["A"], device=dev), "vla_act": c["vla"].to(dev)}, dev)
    loss.backward()
    torch.cuda.synchronize()
    print("get_loss + backward (training forward, dgrad, wgrad_tc, gn_mish_bwd, colsum):", float(loss))
    from vla_touch_b200.lstm_train import LstmLossBackwardProgram
    A, Fd, B, T = 7, 64, 24, 8
    mods = {"force_encoder": syn.synth_state_dict(shp.mlp_shapes([Fd, 128, 128]), 41, "lstm.force_encoder."),
            "lstm": syn.synth_state_dict(shp.lstm_shapes(128 + A), 41, "lstm.lstm."),
            "output_head": syn.synth_state_dict(shp.lstm_head_shapes(256, A), 41, "lstm.output_head.")}
    lp = LstmLossBackwardProgram(mods, A, Fd, B, T, torch.device(dev))
    lp.set_inputs(syn.det_uniform("s.vla", (B, T, A), 1, -1.0, 1.0), syn.det_normal("s.f", (B, T, Fd), 1), syn.det_normal("s.c", (B, 256), 1),
                  syn.det_uniform("s.e", (B, T, A), 1, -1.0, 1.0))
    lp.run()
    torch.cuda.synchronize()
    print("LSTM forward + BPTT (lstm_tc, lstm_bwd_tc):", lp.loss())
    # episode store (batch_gather_kernel, feature cache through a 2-layer DinoV2) and the RDT hand-off (chunk_handoff_kernel)
    import tempfile
    import numpy as np
    from vla_touch_b200 import controller_dataset as cd
    from vla_touch_b200 import episode_store as es
    from vla_touch_b200.rdt_handoff import handoff_action_chunk
    from vla_touch_b200.synthetic import synth_episode
    with tempfile.TemporaryDirectory() as td:
        for e in range(2):
            es.write_episode_shard(synth_episode(e, 30, 224, still_frames=1), os.path.join(td, f"episode_{e}.vtep"))
        ds = cd.ControllerDataset(td, horizon=8, use_images=True, image_size=224)
        store = ds.device_store(dev, image_encoder=ctl.image_encoder, feature_chunk=16)
    got = store.gather(np.arange(len(ds)))
    torch.cuda.synchronize()
    print("episode store gather (batch_gather_kernel):", len(ds), "samples, branch", got["branch"].tolist(), float(got["vla_act"].abs().max()))
    raw, chunk = handoff_action_chunk(torch.randn(2, 64, 128, device=dev).bfloat16(), 32)
    torch.cuda.synchronize()
    print("RDT hand-off (chunk_handoff_kernel):", tuple(raw.shape), tuple(chunk.shape))


if __name__ == "__main__":
    main()
