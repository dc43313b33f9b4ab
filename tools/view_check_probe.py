"""Latency of the agent's view check on the device (check_around, BSCAgent.py:700-731): CLIP ViT-H/14 image tower over the views of
one or two full turns + goal embedding (image or text tower) + view scoring.  CUDA events, warm-up, median of REPS.
Seeded random weights of the real architecture (no checkpoint offline)."""
import importlib
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("bsc-nav_b200")
REPS = int(os.environ.get("REPS", 7))


def timed(fn):
    for _ in range(2):
        fn()
    ms = []
    for _ in range(REPS):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return statistics.median(ms)


def main():
    img_cfg, txt_cfg = pkg.OPENCLIP_H14, pkg.OPENCLIP_H14_TEXT
    img = pkg.ViTEngine(img_cfg, pkg.random_weights(img_cfg, 0), max_batch=24)
    txt = pkg.ViTEngine(txt_cfg, pkg.random_weights(txt_cfg, 1), max_batch=8)
    g = torch.Generator().manual_seed(0)
    out = {}
    ids = torch.zeros(1, txt_cfg.text_context, dtype=torch.int32)
    ids[0, 0], ids[0, 1:6], ids[0, 6] = txt_cfg.vocab - 2, torch.randint(1, 40000, (5,), generator=g), txt_cfg.vocab - 1
    ids = ids.cuda()
    for n_views in (12, 24):
        views = torch.randn(n_views, 3, 224, 224, generator=g).cuda()
        goal_img = torch.randn(1, 3, 224, 224, generator=g).cuda()
        t_img = timed(lambda: img.encode_image(views))
        out[f"encode_image_{n_views}_views_ms"] = round(t_img, 3)
        out[f"encode_image_{n_views}_views_tflops"] = round(n_views * img_cfg.flops_per_frame() / t_img / 1e9, 1)
        out[f"check_image_goal_{n_views}_views_ms"] = round(timed(
            lambda: pkg.score_views(img.encode_image(views), img.encode_image(goal_img)[0])), 3)
        out[f"check_text_goal_{n_views}_views_ms"] = round(timed(
            lambda: pkg.score_views(img.encode_image(views), txt.encode_text(ids)[0])), 3)
    out["encode_text_1_prompt_ms"] = round(timed(lambda: txt.encode_text(ids)), 3)
    out["encode_text_1_prompt_graph_ms"] = round(timed(lambda: txt.encode_text(ids, graph=True)), 3)
    one = torch.randn(1, 3, 224, 224, generator=g).cuda()
    out["encode_image_1_goal_ms"] = round(timed(lambda: img.encode_image(one)), 3)
    out["encode_image_1_goal_graph_ms"] = round(timed(lambda: img.encode_image(one, graph=True)), 3)
    labels = ids.repeat(8, 1)
    out["encode_text_8_labels_ms"] = round(timed(lambda: txt.encode_text(labels)), 3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
