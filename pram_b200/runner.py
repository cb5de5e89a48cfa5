"""Frame-parallel localization runner: SFD2 -> SegNetViT -> GML(+Sinkhorn) -> PnP/RANSAC on a batch of
frames, device-resident end to end (no D2H -> numpy -> H2D round trips between the stages, which the
reference's loop does at every stage boundary, localization/loc_by_rec_online.py:109-189).

One process per GPU; frames are independent units (relocalisation mode, reference README.md:62), so
multi-GPU is plain sharding of frames across ranks with a single gather of fixed-size pose records.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import ops
from .nets.gml import GML
from .nets.segnetvit import SegNetViT
from .nets.sfd2 import ResNet4x


@dataclass
class SyntheticMap:
    """Per-frame reference set for the synthetic stream: the frame's own
    keypoints/descriptors under a seeded permutation, lifted to 3-D with a known pose, 20 % outliers."""
    descriptors: torch.Tensor  # [B,N,128]
    keypoints: torch.Tensor    # [B,N,2]
    xyz: torch.Tensor          # [B,N,3]
    perm: torch.Tensor         # [B,N]  reference j  <-  query perm[j]
    outlier: torch.Tensor      # [B,N] bool
    R: torch.Tensor            # [B,3,3]
    t: torch.Tensor            # [B,3]
    num: Optional[torch.Tensor] = None   # [B] int32: real reference keypoints per set (the rest of the N slots is padding)


class LocalizationPipeline:
    def __init__(self, sfd2: ResNet4x, segnet: SegNetViT, matcher: GML, max_keypoints: int = 1024,
                 focal: float = 525.0, ransac_max_error: float = 8.0, device='cuda', landmarks_per_frame: int = 1):
        self.dev = torch.device(device)
        self.sfd2 = sfd2.to(self.dev)
        self.segnet = segnet.to(self.dev)
        self.matcher = matcher.to(self.dev)
        self.K = max_keypoints
        # > 1: every frame is matched against that many candidate landmarks in ONE batched matcher call ("multi-landmark
        # match_features_batch", BASELINE.json configs[4]; the reference runs these calls one after the other,
        # multimap3d.py:114-239) and the pose with most inliers wins
        self.L = landmarks_per_frame
        self.focal = focal
        self.max_error = ransac_max_error
        self.cfg = {'min_keypoints': 128, 'max_keypoints': max_keypoints}
        self.num_hypotheses = 1024
        self.pre_filtering_th = 0.95  # configs/config_train_7scenes_sfd2.yaml: pre_filtering_th
        self.seg_k = 20               # configs/config_train_7scenes_sfd2.yaml: seg_k
        import os
        self.two_streams = os.environ.get('PRAM_TWO_STREAMS', '1') != '0'
        self._side = None

    # -- stages -------------------------------------------------------------------------------
    def features(self, images: torch.Tensor) -> Dict[str, torch.Tensor]:
        """images [B,3,H,W] normalised -> padded keypoints/descriptors + 256-d recognition features."""
        f = self.sfd2.extract_batched(images, self.cfg)
        f['seg_descriptors'] = ops.sample_features(f['mid_features_nhwc'], f['keypoints'], f['num_keypoints'], 4, False)
        return f

    def recognize(self, f: Dict[str, torch.Tensor], image_shape) -> torch.Tensor:
        """-> landmark logits [B,K,n_class] (reference loc_by_rec_online.py:130)."""
        return self.segnet({'seg_descriptors': f['seg_descriptors'], 'keypoints': f['keypoints'],
                            'num_keypoints': f['num_keypoints'],   # padded slots of the [B, K] layout get no attention
                            'image': torch.empty(image_shape, device='meta')})['prediction']

    def match(self, f: Dict[str, torch.Tensor], smap: SyntheticMap, image_shape) -> Dict[str, torch.Tensor]:
        """query (frame) vs reference (map) sets, one pair per frame (reference singlemap3d.py:143-154,
        including its (1,3,W,H) image_shape convention)."""
        b, _, h, w = image_shape
        shp = (1, 3, w, h)
        d0, k0 = f['descriptors'], f['keypoints']
        if self.L > 1:  # pair (frame b, landmark l) = row b * L + l of the reference set
            d0, k0 = d0.repeat_interleave(self.L, 0), k0.repeat_interleave(self.L, 0)
        # slots j >= num_keypoints[b] of the fixed [B, K] layout are padding (keypoint (0,0), zero descriptor): the counts
        # travel with the batch, so padding gets no attention, is not part of the Sinkhorn problem and comes back unmatched
        # -- every frame is matched exactly as if it were run alone with its n[b] keypoints, like the reference does
        nk = f['num_keypoints'] if self.L == 1 else f['num_keypoints'].repeat_interleave(self.L, 0)
        return self.matcher({'descriptors0': d0, 'keypoints0': k0, 'num_keypoints0': nk,
                             'descriptors1': smap.descriptors, 'keypoints1': smap.keypoints, 'num_keypoints1': smap.num,
                             'image_shape0': shp, 'image_shape1': shp})

    def _recognition(self, f: Dict[str, torch.Tensor], shape, out: Dict[str, torch.Tensor]):
        out['prediction'] = self.recognize(f, shape)
        out['labels'] = out['prediction'].argmax(-1)
        # recognition -> matching glue (reference frame.py:96-121, multimap3d.py:348-379), kept on the device
        b, k, c = out['prediction'].shape
        bg, sid, non_bg = ops.segmentation(out['prediction'].reshape(b * k, c), self.pre_filtering_th)
        valid = torch.arange(k, device=self.dev)[None] < f['num_keypoints'][:, None]     # padded slots vote for no landmark
        out['seg_ids'], out['non_bg'] = sid.view(b, k), non_bg.view(b, k) & valid
        out['landmarks'] = ops.rank_landmarks(out['prediction'], out['non_bg'], self.seg_k, max_ranks=8)

    def localize(self, images: torch.Tensor, smap: Optional[SyntheticMap] = None) -> Dict[str, torch.Tensor]:
        shape = tuple(images.shape)
        f = self.features(images)
        # cand_overflow[b]: NMS left more survivors than the candidate buffer holds (exact plateaus in the score map); the
        # sync-free batched path cannot repeat the frame with a full buffer the way extract_local_global does, so callers
        # check this flag (bench.py does after the timed region; a true entry means that frame's keypoints are a subset)
        out = {'keypoints': f['keypoints'], 'num_keypoints': f['num_keypoints'], 'scores': f['scores'],
               'descriptors': f['descriptors'], 'cand_overflow': f['cand_count'] > f['cand_cap']}
        if smap is None or not self.two_streams:
            self._recognition(f, shape, out)
            if smap is not None:
                m = self.match(f, smap, shape)
                out.update(m)
                out.update(self.pose(f, m, smap, shape))
            return out
        # recognition (SegNetViT + ranking) and matching + pose (GML, Sinkhorn, RANSAC) only share the features: they
        # run on two streams (a fork/join inside the captured graph), so the prologue / tail of every persistent
        # kernel of one branch is filled by the other branch
        main = torch.cuda.current_stream(self.dev)
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.dev)
        side = self._side
        side.wait_stream(main)
        with torch.cuda.stream(side):
            self._recognition(f, shape, out)
        m = self.match(f, smap, shape)
        out.update(m)
        out.update(self.pose(f, m, smap, shape))
        main.wait_stream(side)
        for v in (out['prediction'], out['labels'], out['seg_ids'], out['non_bg'], *out['landmarks'].values()):
            v.record_stream(main)
        return out

    def pose(self, f: Dict[str, torch.Tensor], m: Dict[str, torch.Tensor], smap: SyntheticMap, image_shape):
        """2D-3D matches -> absolute pose per frame (reference singlemap3d.py:156-175: matched keypoints
        + 0.5, matched xyz, ransac max_error from the config), entirely on the device."""
        h, w = image_shape[-2:]
        kp = f['keypoints'] if self.L == 1 else f['keypoints'].repeat_interleave(self.L, 0)
        r = ops.ransac_pnp(kp, m['matches0'], smap.xyz, self.focal, self.focal, w / 2.0, h / 2.0,
                           self.max_error, pixel_shift=0.5, num_hypotheses=self.num_hypotheses, seed=0)
        if self.L > 1:  # per frame: the landmark whose pose has most inliers (ties: the first, like the reference's loop order)
            b = f['keypoints'].shape[0]
            best = r['num_inliers'].view(b, self.L).argmax(1)
            rows = torch.arange(b, device=self.dev) * self.L + best
            r = {k: v[rows] for k, v in r.items()}
            r['best_landmark'] = best
        out = {'qvec': r['qvec'], 'tvec': r['tvec'], 'num_inliers': r['num_inliers'], 'inliers': r['inliers'],
               'pose_success': r['success']}
        if self.L > 1:
            out['best_landmark'] = r['best_landmark']
        return out

    # -- per-stage device timers (reference loc_by_rec_online.py:108-134, 212-222) ---------------------
    @torch.no_grad()
    def localize_timed(self, images: torch.Tensor, smap: Optional[SyntheticMap] = None):
        """``localize`` run eagerly on ONE stream with CUDA events between the stages.  Returns (out, times) with
        ``times`` = {'time_feat', 'time_rec', 'time_loc', 'time_total'} in SECONDS for the whole batch -- the quantities the
        reference stores on every Frame (``Frame.time_feat / time_rec / time_loc``; ``time_ref`` = 0: no refinement round in
        this path) -- measured on the device instead of with time.time() around asynchronous launches."""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        shape = tuple(images.shape)
        ev[0].record()
        f = self.features(images)
        ev[1].record()
        out = {'keypoints': f['keypoints'], 'num_keypoints': f['num_keypoints'], 'scores': f['scores'],
               'descriptors': f['descriptors'], 'cand_overflow': f['cand_count'] > f['cand_cap']}
        self._recognition(f, shape, out)
        ev[2].record()
        if smap is not None:
            m = self.match(f, smap, shape)
            out.update(m)
            out.update(self.pose(f, m, smap, shape))
        ev[3].record()
        torch.cuda.synchronize(self.dev)
        t = {'time_feat': ev[0].elapsed_time(ev[1]) * 1e-3, 'time_rec': ev[1].elapsed_time(ev[2]) * 1e-3,
             'time_loc': ev[2].elapsed_time(ev[3]) * 1e-3, 'time_ref': 0.0}
        t['time_total'] = t['time_feat'] + t['time_rec'] + t['time_loc'] + t['time_ref']
        return out, t

    # -- CUDA graph: one replay per batch instead of ~1000 launches + tensor-map encodes -----------------
    @torch.no_grad()
    def capture(self, images: torch.Tensor, smap: Optional[SyntheticMap] = None):
        """Capture ``localize`` for this batch shape into a CUDA graph (all shapes are static and no stage
        syncs with the host).  Returns the number of library kernels inside one replay."""
        from . import _lib
        self._static_in = images.clone()
        # small batches are bound by the latency of ~150 dependent launches: let every persistent kernel release its
        # successor as soon as its own CTAs are resident (programmatic dependent launch, early trigger; baked into the
        # captured nodes).  At batch 32 the early trigger costs 0.6 %, so it stays off there.
        lib = _lib.load()
        pdl_saved = lib.pram_get_pdl()
        if pdl_saved == 1 and images.shape[0] <= 4:
            lib.pram_set_pdl(2)
        try:
            side = torch.cuda.Stream(device=self.dev)
            side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(side):
                for _ in range(2):
                    self.localize(self._static_in, smap)
            torch.cuda.current_stream(self.dev).wait_stream(side)
            torch.cuda.synchronize(self.dev)
            self._graph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(self._graph):
                self._static_out = self.localize(self._static_in, smap)
            self.graph_launches = _lib.launch_count() - n0
        finally:
            lib.pram_set_pdl(pdl_saved)
        return self.graph_launches

    def replay(self, images: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        if images is not None:
            self._static_in.copy_(images, non_blocking=True)
        self._graph.replay()
        return self._static_out

    # -- synthetic map ------------------------------------------------------------------------
    @torch.no_grad()
    def build_synthetic_map(self, images: torch.Tensor, seed: int = 0, outlier_frac: float = 0.2,
                            ref_keypoints: Optional[int] = None) -> SyntheticMap:
        """One reference set per (frame, landmark): the frame's own keypoints under a seeded permutation (the first
        ``ref_keypoints`` of it when given: BASELINE.json configs[4] matches 4096 query keypoints against 1024 reference
        keypoints per landmark), lifted to 3-D with a known pose, ``outlier_frac`` replaced by outliers."""
        f = self.features(images)
        nkp = f['num_keypoints'].repeat_interleave(self.L, 0).cpu()
        f = {'keypoints': f['keypoints'].repeat_interleave(self.L, 0), 'descriptors': f['descriptors'].repeat_interleave(self.L, 0)}
        b, k, _ = f['keypoints'].shape
        nb = b // self.L
        g = torch.Generator(device='cpu').manual_seed(seed)
        perms = []
        for i in range(b):
            pm = torch.randperm(k, generator=g)
            n_i = int(nkp[i])
            if n_i < k:  # frame with fewer keypoints than slots: its real keypoints first (same relative order), padding last
                pm = torch.cat([pm[pm < n_i], pm[pm >= n_i]])
            perms.append(pm)
        perm = torch.stack(perms).to(self.dev)
        num = nkp.clamp(max=k).to(torch.int32)
        if ref_keypoints is not None and ref_keypoints < k:
            perm = perm[:, :ref_keypoints].contiguous()
            k = ref_keypoints
            num = num.clamp(max=k)
        kp = torch.gather(f['keypoints'], 1, perm[..., None].expand(-1, -1, 2))
        desc = torch.gather(f['descriptors'], 1, perm[..., None].expand(-1, -1, 128)).clone()
        # known pose per frame: small rotation (<= 15 deg) and translation (<= 0.5 m)
        # one camera pose per FRAME (its landmarks share the world frame)
        ax = torch.nn.functional.normalize(torch.randn(nb, 3, generator=g), dim=-1).repeat_interleave(self.L, 0)
        ang = (torch.rand(nb, generator=g) * 15.0 * 3.14159265 / 180.0).repeat_interleave(self.L, 0)
        Kx = torch.zeros(b, 3, 3)
        Kx[:, 0, 1], Kx[:, 0, 2], Kx[:, 1, 0] = -ax[:, 2], ax[:, 1], ax[:, 2]
        Kx[:, 1, 2], Kx[:, 2, 0], Kx[:, 2, 1] = -ax[:, 0], -ax[:, 1], ax[:, 0]
        R = torch.eye(3)[None] + torch.sin(ang)[:, None, None] * Kx + (1 - torch.cos(ang))[:, None, None] * (Kx @ Kx)
        t = (torch.rand(nb, 3, generator=g) - 0.5).repeat_interleave(self.L, 0)
        z = 1.0 + 4.0 * torch.rand(b, k, generator=g)
        outl = torch.rand(b, k, generator=g) < outlier_frac
        R, t, z, outl = R.to(self.dev), t.to(self.dev), z.to(self.dev), outl.to(self.dev)
        h, w = images.shape[-2:]
        cx, cy = w / 2.0, h / 2.0
        ray = torch.stack([(kp[..., 0] + 0.5 - cx) / self.focal, (kp[..., 1] + 0.5 - cy) / self.focal,
                           torch.ones_like(z)], -1) * z[..., None]
        xyz = torch.einsum('bji,bnj->bni', R, ray - t[:, None])  # X = R^T (x_cam - t)
        rnd_desc = torch.nn.functional.normalize(torch.randn(b, k, 128, generator=g), dim=-1).to(self.dev)
        rnd_xyz = (torch.rand(b, k, 3, generator=g) * 4 - 2).to(self.dev)
        desc = torch.where(outl[..., None], rnd_desc, desc)
        xyz = torch.where(outl[..., None], rnd_xyz, xyz)
        return SyntheticMap(desc.contiguous(), kp.contiguous(), xyz.contiguous(), perm, outl, R, t, num.to(self.dev))


# ---- frame sharding across ranks (one process per GPU) ---------------------------------------------------

def shard_frames(n_frames: int, rank: int, world: int):
    """Frame ids owned by ``rank``: round-robin, ``i -> rank i % world``.  Frames are
    independent units, so this is the whole multi-GPU data path; no activation ever crosses ranks."""
    return list(range(rank, n_frames, world))


def pack_pose_records(frame_ids, qvec: torch.Tensor, tvec: torch.Tensor, num_inliers: torch.Tensor) -> torch.Tensor:
    """[n, 9] float64 records ``[frame_id, qw,qx,qy,qz, tx,ty,tz, num_inliers]`` (72 bytes per frame)."""
    rec = torch.zeros((len(frame_ids), 9), dtype=torch.float64, device=qvec.device)
    rec[:, 0] = torch.as_tensor(frame_ids, dtype=torch.float64, device=qvec.device)
    rec[:, 1:5] = qvec.double()
    rec[:, 5:8] = tvec.double()
    rec[:, 8] = num_inliers.double()
    return rec


def gather_pose_records(rec: torch.Tensor) -> torch.Tensor:
    """The path's single collective: all_gather of the fixed-size pose records (NCCL over NVLink on GPUs,
    gloo in the CPU tests), returned sorted by frame id.  Every rank must contribute the same number of
    records (pad with frame_id = -1)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rec
    out = [torch.zeros_like(rec) for _ in range(dist.get_world_size())]
    dist.all_gather(out, rec)
    allrec = torch.cat(out, 0)
    allrec = allrec[allrec[:, 0] >= 0]
    return allrec[torch.argsort(allrec[:, 0])]
