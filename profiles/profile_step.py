"""One adversarial step of the bench workload between cudaProfilerStart/Stop (for ncu --profile-from-start off).

    python profiles/profile_step.py --mode tf32 [--what step|decode]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="tf32")
    ap.add_argument("--what", default="step")
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    import gic_b200
    from gic_b200.args import default_args
    from gic_b200.training import GANInstructor
    gic_b200.set_gemm_mode(bench.MODES[a.mode])
    cfg = bench.WORKLOADS[a.workload]
    B, L, V = cfg["B"], cfg["L"], cfg["V"]
    dev = torch.device("cuda:0")
    args = default_args(vocab_size=V, gen_embed_dim=cfg["E"], gen_hidden_dim=cfg["H"], gen_num_layers=cfg["layers"],
                        disc_num_filters=list(cfg["filters"]), conditional_gan=1, feature_dim=cfg["feat"], device="cuda")
    torch.manual_seed(1008)
    inst = GANInstructor(args, device=dev)
    inst.gen.train(); inst.disc.train()
    inst.gen.decoder.temperature = 1.0
    g = torch.Generator(device=dev).manual_seed(1)
    caps = torch.randint(4, V, (B, L), generator=g, device=dev)
    pooled = torch.randn(B, cfg["feat"], generator=g, device=dev)
    u = torch.rand(L, B, V, generator=g, device=dev)
    keep = (torch.rand(3, B * 64, sum(cfg["filters"]), generator=g, device=dev) >= 0.2).to(torch.uint8)

    def run():
        if a.what == "decode":
            with torch.no_grad():
                f = inst.gen.encoder(pooled)
                inst.gen.decoder.sample(f, max_caption_len=L, u=u)
        else:
            inst.adv_step(caps, pooled=pooled, u=u, keep=keep)

    for _ in range(a.warmup):
        run()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    run()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled one", a.what, "in mode", a.mode)


if __name__ == "__main__":
    main()
