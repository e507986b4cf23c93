"""
End-to-end run on an analytic scene (no dataset files are available offline): ray-traced sphere images
from a ring of cameras -> sample-mode training batches built on the device -> NeRF.fit -> held-out views
rendered and scored with the eval.py PSNR. Everything numeric runs in the sm_100a kernels.

    python examples/train_spheres.py [steps] [H]     # defaults: 3000 steps, 100x100 images
"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import nerf_tf2_b200 as nb
from nerf_tf2_b200 import render, scene


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    H = W = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    sc = scene.SyntheticScene(H, W, num_cameras=40, inclination=55.0)
    imgs = np.stack([scene.render_spheres(H, W, sc.K, p) for p in sc.poses])
    test_ids = list(range(0, len(sc), 8))
    train_ids = [i for i in range(len(sc)) if i not in test_ids]
    params = nb.make_params({"system": {"white_bg": True}})
    nerf = nb.setup_model(params, precision="bf16", train_precision="bf16", seed=0)
    train = nb.SampleModeDataset(imgs[train_ids], np.stack([sc.poses[i] for i in train_ids]), sc.bounds, sc.K,
                                 batch_size=4096, seed=1)
    test_poses = [sc.poses[i] for i in test_ids]

    def score():
        return render.evaluate_views(nerf, H, W, test_poses, sc.bounds, sc.K, imgs[test_ids])["mean_psnr"]

    log = [{"step": 0, "test_psnr": score()}]
    done, t_train = 0, 0.0
    while done < steps:
        n = min(500, steps - done)
        torch.cuda.synchronize()
        t0 = time.time()
        hist = nerf.fit(train, epochs=1, steps_per_epoch=n)
        torch.cuda.synchronize()
        t_train += time.time() - t0
        done += n
        log.append({"step": done, "train_psnr_metric": hist.history["psnr_metric"][-1], "loss": hist.history["loss"][-1],
                    "test_psnr": score()})
        print(json.dumps(log[-1]), flush=True)
    print(json.dumps({"steps": steps, "image": [H, W], "train_views": len(train_ids), "test_views": len(test_ids),
                      "train_seconds": round(t_train, 2), "steps_per_s": round(steps / t_train, 1),
                      "test_psnr_start": log[0]["test_psnr"], "test_psnr_end": log[-1]["test_psnr"]}))


if __name__ == "__main__":
    main()
