"""Generate tests/golden/*.npz from the REFERENCE's own modules (build container only).

    python oracle/make_golden.py

The reference ships no tests or golden vectors, so these fixtures are produced
by importing its modules from /root/reference (unmodified, CPU, fp32) on seeded synthetic inputs.
They pin the oracle (tests/test_oracle_pin.py) and, on the GPU box where the reference is absent, the
CUDA path (tests/test_gpu_*.py).  Small by construction (a few hundred KB each).
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import pram_oracle as O, ref_loader as RL  # noqa: E402

OUT = ROOT / 'tests' / 'golden'


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    ref = RL.import_reference()
    torch.manual_seed(0)
    torch.set_num_threads(8)

    # ---- SFD2: shipped weights, 160x120 polygon frame, K=256 -------------------------------
    sd = RL.load_sfd2_state()
    net = ref.sfd2.ResNet4x()
    net.load_state_dict(sd, strict=True)
    net.eval()
    img = O.frame_tensor(120, 160, seed=3)
    cfg = {'min_keypoints': 32, 'max_keypoints': 64}
    with torch.no_grad():
        r = net.extract_local_global({'image': img}, cfg)
        k = r['keypoints'][0]
        sc64, seg64 = net.sample(r['score_map'], r['mid_features'], k, norm_desc=False)
        r_all = net.extract_local_global({'image': img}, {'min_keypoints': 32, 'max_keypoints': 4096})
        k = r_all['keypoints'][0]  # 119 keypoints, row-major: used for recognition / matching below
        sc, seg = net.sample(r['score_map'], r['mid_features'], k, norm_desc=False)
        nms4 = ref.sfd2.simple_nms(r['score_map'], 4)
        nms3 = ref.sfd2.simple_nms(r['score_map'], 3)
    np.savez_compressed(OUT / 'sfd2_160x120.npz', image=img.numpy(), score_map=r['score_map'].numpy(),
                        keypoints=r['keypoints'][0].numpy(), scores=r['scores'][0].numpy(),
                        descriptors=r['descriptors'][0].numpy(), sample_scores=sc64.numpy(),
                        seg_descriptors=seg64.numpy().astype(np.float32),
                        descriptors_all=r_all['descriptors'][0].numpy(), nms4=nms4.numpy(), nms3=nms3.numpy(),
                        keypoints_all=r_all['keypoints'][0].numpy(), scores_all=r_all['scores'][0].numpy(),
                        max_keypoints=64, min_keypoints=32)

    # ---- SegNetViT: seeded random state (no shipped recognition weights) ---------------------
    ssd = RL.random_segnetvit_state(n_class=113, seed=0)
    m = ref.segnetvit.SegNetViT({'n_class': 113, 'n_layers': 15, 'output_dim': 1024, 'descriptor_dim': 256})
    m.load_state_dict(ssd, strict=True)
    m.eval()
    with torch.no_grad():
        pred = m({'seg_descriptors': seg.t()[None], 'keypoints': k[None], 'image': img})['prediction']
    np.savez_compressed(OUT / 'segnetvit_seed0.npz', seg_descriptors=seg.t().numpy(), keypoints=k.numpy(),
                        image_shape=np.array(img.shape), prediction=pred[0].numpy(), n_class=113, seed=0)

    # ---- GML: shipped weights, self-match under a permutation --------------------------------
    gsd = RL.load_gml_state()
    g = ref.gml.GML({})
    g.load_state_dict(gsd, strict=True)
    g.eval()
    gen = torch.Generator().manual_seed(1)
    perm = torch.randperm(k.shape[0], generator=gen)[:100]  # N != M on purpose
    d0 = r_all['descriptors'][0].t()[None]
    data = {'descriptors0': d0, 'descriptors1': d0[:, perm], 'keypoints0': k[None], 'keypoints1': k[perm][None],
            'image_shape0': (1, 3, 160, 120), 'image_shape1': (1, 3, 160, 120)}
    with torch.no_grad():
        out = g(data)
    np.savez_compressed(OUT / 'gml_selfmatch.npz', descriptors0=d0[0].numpy(), keypoints0=k.numpy(),
                        perm=perm.numpy(), matches0=out['matches0'][0].numpy(), matches1=out['matches1'][0].numpy(),
                        scores0=out['matching_scores0'][0].numpy(), scores1=out['matching_scores1'][0].numpy())

    # ---- Sinkhorn + matches on a random distance matrix --------------------------------------
    gen = torch.Generator().manual_seed(2)
    dist = torch.randn(2, 70, 93, generator=gen) * 3
    for i in range(40):  # plant some confident correspondences
        dist[:, i, (i * 7) % 93] += 12
    bin_score = torch.tensor(1.3)
    with torch.no_grad():
        P = ref.gml.sink_algorithm(dist, bin_score, 20)
        i0, i1, s0, s1 = g.compute_matches(P, 0.2)
    np.savez_compressed(OUT / 'sinkhorn_70x93.npz', dist=dist.numpy(), bin_score=bin_score.numpy(), P=P.numpy(),
                        matches0=i0.numpy(), matches1=i1.numpy(), scores0=s0.numpy(), scores1=s1.numpy())
    for f in sorted(OUT.glob('*.npz')):
        print(f.name, f.stat().st_size // 1024, 'KB')


if __name__ == '__main__':
    main()
