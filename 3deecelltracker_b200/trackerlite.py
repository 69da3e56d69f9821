"""TrackerLite and its EM operators on the GPU -- drop-in for CellTracker/trackerlite.py (hot path).

Class API: TrackerLite.predict_cell_positions / predict_cell_positions_ensemble / match_by_ffn
(trackerlite.py:33-142).  Operators: simple_match (:242), prgls_quick (:262), prgls_with_two_ref (:309),
get_volumes_list / evenly_distributed_volumes (:420-438).
"""
from pathlib import Path
from typing import List

import numpy as np
import torch

from . import _lib
from ._device import WORKSPACE, aligned_ptr, stream_ptr, to_device
from .coord_image_transformer import Coordinates
from .ffn import FFN, initial_matching_ffn, normalize_points
from .track import MODE_LITE, EmProblem, run_em, trim_mean_device

FIGURE = "figure"
COORDS_REAL = "coords_real"
LABELS = "labels"
TRACK_RESULTS = "track_results"
SEG = "seg"

BETA, LAMBDA, MAX_ITERATION = (3, 3, 2000)     # trackerlite.py:29
K_POINTS = 20                                  # trackerlite.py:30


def simple_match(initial_match_matrix, threshold=0.1):
    """Greedy one-to-one matching (trackerlite.py:242-259).
    Returns (normalized_prob (M,N) in the input dtype, pairs_px2 [(ref, tgt)])."""
    corr = np.asarray(initial_match_matrix)
    f64 = corr.dtype != np.float32
    dev = to_device(corr, torch.float64 if f64 else torch.float32)
    m, n = (int(s) for s in dev.shape)
    prior = torch.empty((m, n), dtype=torch.float64, device=dev.device)
    pairs = torch.zeros((min(m, n) + 1, 2), dtype=torch.int32, device=dev.device)
    n_pairs = torch.zeros(1, dtype=torch.int32, device=dev.device)
    lib = _lib.lib()
    ws = WORKSPACE.get("greedy", lib.ct_greedy_workspace_bytes(n, m))
    wp = aligned_ptr(ws)
    _lib.check(lib.ct_greedy_prior(dev.data_ptr(), 1 if f64 else 0, m, n, MODE_LITE, float(threshold),
                                   prior.data_ptr(), pairs.data_ptr(), n_pairs.data_ptr(), wp,
                                   ws.numel() - (wp - ws.data_ptr()), stream_ptr()))
    k = int(n_pairs.item())
    pr = pairs[:k].cpu().numpy()
    pairs_px2 = np.stack([pr[:, 1], pr[:, 0]], axis=1) if k else np.array([])
    return prior.cpu().numpy().astype(corr.dtype if corr.dtype in (np.float32, np.float64) else np.float64), pairs_px2


def prgls_with_two_ref(init_match_mxn, ptrs_tgt_mx3, prts_ref_nx3, tracked_ref_lx3, beta, lambda_,
                       max_iteration=MAX_ITERATION):
    """trackerlite.py:309-358.  Returns (predicted tracked (L,3), posterior (M,N)) float64 ndarrays."""
    p = EmProblem(np.asarray(prts_ref_nx3, dtype=np.float64), np.asarray(ptrs_tgt_mx3, dtype=np.float64),
                  init_match_mxn, tracked=np.asarray(tracked_ref_lx3, dtype=np.float64), prior_given=True)
    run_em([p], MODE_LITE, beta, lambda_, max_iteration, 1.0, 0.1)
    return p.tracked_out.cpu().numpy(), p.post.cpu().numpy()


def prgls_quick(init_match_mxn, ptrs_tgt_mx3, tracked_ref_nx3, beta, lambda_, max_iteration=MAX_ITERATION):
    """trackerlite.py:262-306 (the reference set moves itself)."""
    return prgls_with_two_ref(init_match_mxn, ptrs_tgt_mx3, tracked_ref_nx3, tracked_ref_nx3, beta, lambda_,
                              max_iteration)


def evenly_distributed_volumes(current_vol, sampling_number, start_vol=1):
    """trackerlite.py:420-424."""
    interval = (current_vol - start_vol) // sampling_number
    start = (current_vol - start_vol) % sampling_number + start_vol
    return list(range(start, current_vol - interval + 1, interval))


def get_volumes_list(current_vol, skip_volumes, sampling_number=20, adjacent=False, start_vol=1):
    """trackerlite.py:427-438."""
    assert current_vol > start_vol, f"current_vol (={current_vol}) should be larger than start_vol (={start_vol})"
    if current_vol - start_vol < sampling_number:
        vols_list = list(range(start_vol, current_vol))
    elif adjacent:
        vols_list = list(range(current_vol - sampling_number, current_vol))
    else:
        vols_list = evenly_distributed_volumes(current_vol, sampling_number, start_vol=start_vol)
    return [vol for vol in vols_list if vol not in skip_volumes]


class TrackerLite:
    """Tracks cells in 3D time-lapse images from per-volume segmented coordinates with a trained FFN
    (trackerlite.py:33-142).  Results directory layout is the reference's:
    <results_dir>/seg/coords%06d.npy in, <results_dir>/track_results/coords_real/coords%06d.npy reused for
    ensembles."""

    def __init__(self, results_dir, ffn_model_name, proofed_coords_vol1, miss_frame=None, basedir="ffn_models"):
        if miss_frame is not None and not isinstance(miss_frame, List):
            raise TypeError(f"miss_frame should be a list or None, but got {type(miss_frame)}")
        self.results_dir = Path(results_dir)
        (self.results_dir / TRACK_RESULTS / FIGURE).mkdir(parents=True, exist_ok=True)
        (self.results_dir / TRACK_RESULTS / COORDS_REAL).mkdir(parents=True, exist_ok=True)
        (self.results_dir / TRACK_RESULTS / LABELS).mkdir(parents=True, exist_ok=True)
        if isinstance(ffn_model_name, FFN):
            self.ffn_model_path = None
            self.ffn_model = ffn_model_name
        else:
            self.ffn_model_path = Path(basedir) / (ffn_model_name + ".npz")
            self.ffn_model = FFN()
            self.ffn_model.load_weights(str(self.ffn_model_path))      # raises ValueError like trackerlite.py:64-65
        self.proofed_coords_vol1 = proofed_coords_vol1
        self.miss_frame = [] if miss_frame is None else miss_frame

    def _get_segmented_pos(self, t):
        return Coordinates(np.load(str(self.results_dir / SEG / f"coords{str(t).zfill(6)}.npy")),
                           interpolation_factor=self.proofed_coords_vol1.interpolation_factor,
                           voxel_size=self.proofed_coords_vol1.voxel_size, dtype="raw")

    def _predict_device_batch(self, members, seg_t2_real, beta, lambda_):
        """The device pipeline of predict_cell_positions for several reference volumes at once: per member
        (segmented points of t1, confirmed points of t1) an FFN match, then ONE batched EM launch (simple_match +
        prgls_with_two_ref, one persistent CTA per member) with no host round trip in between.
        Returns [(EmProblem with outputs on the device, mean_t1, scale_t1)]."""
        probs, paras = [], []
        for seg_t1_real, confirmed_t1_real in members:
            conf_norm, (mean_t1, scale_t1) = normalize_points(confirmed_t1_real, return_para=True)
            seg2_norm = (seg_t2_real - mean_t1) / scale_t1
            seg1_norm = (seg_t1_real - mean_t1) / scale_t1
            ref_dev = to_device(seg1_norm.astype(np.float64), torch.float64)
            tgt_dev = to_device(seg2_norm.astype(np.float64), torch.float64)
            corr = self.ffn_model.match_device(ref_dev, tgt_dev, K_POINTS)
            probs.append(EmProblem(ref_dev, tgt_dev, corr, tracked=conf_norm.astype(np.float64), prior_given=False))
            paras.append((mean_t1, scale_t1))
        run_em(probs, MODE_LITE, beta, lambda_, MAX_ITERATION, 1.0, 0.1)
        return [(p, m, sc) for p, (m, sc) in zip(probs, paras)]

    def _predict_device(self, seg_t1_real, seg_t2_real, confirmed_t1_real, beta, lambda_):
        """One member of `_predict_device_batch`: FFN match -> simple_match -> prgls_with_two_ref."""
        return self._predict_device_batch([(seg_t1_real, confirmed_t1_real)], seg_t2_real, beta, lambda_)[0]

    def predict_cell_positions(self, t1, t2, confirmed_coord_t1=None, beta=BETA, lambda_=LAMBDA, draw_fig=False):
        """Positions of the confirmed cells of t1 at t2 (trackerlite.py:70-109)."""
        assert t2 not in self.miss_frame
        segmented_pos_t1 = self._get_segmented_pos(t1)
        segmented_pos_t2 = self._get_segmented_pos(t2)
        if confirmed_coord_t1 is None:
            confirmed_coord_t1 = segmented_pos_t1
        prob, mean_t1, scale_t1 = self._predict_device(segmented_pos_t1.real, segmented_pos_t2.real,
                                                       confirmed_coord_t1.real, beta, lambda_)
        tracked_coords_t2 = prob.tracked_out.cpu().numpy() * scale_t1 + mean_t1
        return Coordinates(tracked_coords_t2, interpolation_factor=self.proofed_coords_vol1.interpolation_factor,
                           voxel_size=self.proofed_coords_vol1.voxel_size, dtype="real")

    def predict_cell_positions_ensemble(self, skipped_volumes, t2, coord_t1, beta, lambda_, sampling_number=20,
                                        adjacent=False, t_start=1):
        """trackerlite.py:111-125: one prediction per reference volume, 10 %-trimmed mean.  The reference calls
        predict_cell_positions once per volume; here all members share one batched EM launch (results are those of
        the per-member calls bit for bit, see tests)."""
        assert t2 not in self.miss_frame
        factor, voxel = self.proofed_coords_vol1.interpolation_factor, self.proofed_coords_vol1.voxel_size
        segmented_pos_t2 = self._get_segmented_pos(t2)
        members = []
        for t1 in get_volumes_list(current_vol=t2, skip_volumes=skipped_volumes, sampling_number=sampling_number,
                                   adjacent=adjacent, start_vol=t_start):
            loaded = np.load(str(self.results_dir / TRACK_RESULTS / COORDS_REAL / f"coords{str(t1).zfill(6)}.npy"))
            loaded_ = Coordinates(loaded, coord_t1.interpolation_factor, coord_t1.voxel_size, dtype="real")
            members.append((self._get_segmented_pos(t1).real, loaded_.real))
        coord_prgls = []
        for prob, mean_t1, scale_t1 in self._predict_device_batch(members, segmented_pos_t2.real, beta, lambda_):
            tracked = prob.tracked_out.cpu().numpy() * scale_t1 + mean_t1
            coord_prgls.append(Coordinates(tracked, factor, voxel, dtype="real").real)      # float32 round trip of :109
        stack = to_device(np.asarray(coord_prgls, dtype=np.float64), torch.float64)
        mean = trim_mean_device(stack, 0.1).cpu().numpy()
        return Coordinates(mean, interpolation_factor=factor, voxel_size=voxel, dtype="real")

    def match_by_ffn(self, t1, t2, confirmed_coord_t1=None):
        """trackerlite.py:127-142 without the plot: returns (matching_matrix, pairs_px2)."""
        assert t2 not in self.miss_frame
        segmented_pos_t1 = self._get_segmented_pos(t1)
        segmented_pos_t2 = self._get_segmented_pos(t2)
        if confirmed_coord_t1 is None:
            confirmed_coord_t1 = segmented_pos_t1
        conf_norm, (mean_t1, scale_t1) = normalize_points(confirmed_coord_t1.real, return_para=True)
        seg2_norm = (segmented_pos_t2.real - mean_t1) / scale_t1
        matching_matrix = initial_matching_ffn(self.ffn_model, conf_norm, seg2_norm, K_POINTS)
        _, pairs_px2 = simple_match(matching_matrix)
        return matching_matrix, pairs_px2
