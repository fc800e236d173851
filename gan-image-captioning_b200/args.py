"""Hyper-parameter surface of the reference (src/args.py:6-280): same flag names, types and defaults.

``get_args(argv)`` parses like the reference; unlike the reference it has no side effects unless
``make_dirs=True`` (the reference mkdirs ``save_dir``/``model_dir`` and auto-increments the experiment
name, src/args.py:259-273)."""
from __future__ import annotations

import argparse
import os


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser()
    m = p.add_argument_group("model")
    m.add_argument("--gen-hidden-dim", type=int, default=512)
    m.add_argument("--gen-embed-dim", type=int, default=32)
    m.add_argument("--gen-num-layers", type=int, default=1)
    m.add_argument("--gen-init", type=str, default="uniform")
    m.add_argument("--disc-embed-dim", type=int, default=64)
    m.add_argument("--disc-num-rep", type=int, default=64)
    # Q10: the reference declares these with type=list (a CLI string would be split into characters);
    # a comma-separated list of ints is accepted here, the defaults are the reference's.
    m.add_argument("--disc-filter-sizes", type=lambda s: [int(x) for x in s.split(",")], default=[3, 4, 5])
    m.add_argument("--disc-num-filters", type=lambda s: [int(x) for x in s.split(",")], default=[300, 300, 300])
    m.add_argument("--disc-init", type=str, default="uniform")
    m.add_argument("--conditional-gan", type=int, default=0)
    d = p.add_argument_group("data")
    d.add_argument("--vocab-size", type=int, default=-1)
    d.add_argument("--max-seq-len", type=int, default=34)
    d.add_argument("--padding-idx", type=int, default=0)
    d.add_argument("--image-size", type=int, default=256)
    d.add_argument("--captions-per-image", type=int, default=1)
    d.add_argument("--dataset_percent", type=float, default=1.0)
    t = p.add_argument_group("training")
    t.add_argument("--pretrain-lr", type=float, default=1e-2)
    t.add_argument("--pretrain-epochs", type=int, default=0)
    t.add_argument("--pre-train-batch-size", type=int, default=64)
    t.add_argument("--pre-eval-batch-size", type=int, default=64)
    t.add_argument("--gen-lr", type=float, default=1e-4)
    t.add_argument("--disc-lr", type=float, default=1e-4)
    t.add_argument("--disc-train-freq", type=int, default=1)
    t.add_argument("--adv-epochs", type=int, default=30)
    t.add_argument("--adv-train-batch-size", type=int, default=64)
    t.add_argument("--adv-eval-batch-size", type=int, default=64)
    t.add_argument("--adv-loss-type", type=str, default="standard")
    t.add_argument("--temperature", type=int, default=100)
    t.add_argument("--temp-adpt", type=str, default="exp")
    t.add_argument("--clip-norm", type=float, default=5.0)
    r = p.add_argument_group("run")
    r.add_argument("--device", type=str, default="cuda")
    r.add_argument("--device-ids", type=int, default=0)
    r.add_argument("--expt-name", type=str, default="debug")
    r.add_argument("--model-dir", type=str, default="models")
    r.add_argument("--data-dir", type=str, default="./data")
    r.add_argument("--save-dir", type=str, default="./save")
    r.add_argument("--adv-log-step", type=int, default=1)
    r.add_argument("--pre-log-step", type=int, default=1)
    r.add_argument("--test-log-step", type=int, default=1)
    r.add_argument("--log-file", type=str, default="log")
    # extension (not in the reference): width of the pooled CNN feature fed to Encoder.linear
    r.add_argument("--feature-dim", type=int, default=512)
    # extension (north-star attention cell): additive attention over a [P, feature-channels] grid
    r.add_argument("--gen-attention", type=int, default=0)
    r.add_argument("--attn-dim", type=int, default=256)
    r.add_argument("--feature-channels", type=int, default=2048)
    return p


def get_args(argv=None, make_dirs: bool = False):
    args = build_parser().parse_args(argv)
    if make_dirs:
        os.makedirs(os.path.join(args.save_dir, args.expt_name, args.model_dir), exist_ok=True)
    return args


def default_args(**over):
    a = build_parser().parse_args([])
    for k, v in over.items():
        if not hasattr(a, k):
            raise AttributeError("unknown hyper-parameter %r" % k)
        setattr(a, k, v)
    return a
