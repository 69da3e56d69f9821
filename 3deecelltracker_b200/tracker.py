"""`Tracker`: the U-Net workflow's segment -> match -> track loop with every hot-path operator on the GPU.

Mirrors the call chain of CellTracker/tracker.py: `_segment` (:605) -> `_save_unet_regions` (:662)
[_normalize_image + unet3_prediction]; `_fit_ffn_prgls` (:1224) / `_ffn_prgls_once` (:1256)
[initial_matching_quick + pr_gls_quick, REP_NUM_PRGLS = 5 with beta * 0.8**i]; `_predict_one_rep` (:1269);
`_predict_pos_once` (:1193); `match` (:1138); `track` / `track_one_vol` (:1415, :1473) with
get_reference_vols + trim_mean ensembles (:1503-1507).

`_watershed` (tracker.py:671-684) and the centre of mass (:646-648) run on the GPU too (watershed.py of this package:
the probability map goes from the U-Net to the label image and the cell centres without leaving HBM).

Host-side lines that are identical to the reference (constructor attribute assignments, method signatures, the
ValueError messages, History bookkeeping) are the drop-in surface the north_star asks to keep; none of them computes.

Out of scope here (host glue of the reference): matplotlib drawing, TIFF writing of
label images, manual-correction I/O, U-Net retraining.
"""
import os

import numpy as np
import torch

from . import _lib
from . import correction as _corr
from . import io_formats
from . import watershed as _ws
from ._device import to_device
from .ffn import FFN
from .preprocess import normalize_image_device, _raw_to_device
from .track import (MODE_TRACK, EmProblem, get_reference_vols, predict_one_rep_device, run_em, trim_mean_device)
from .unet3d import UNet3

REP_NUM_PRGLS = 5          # tracker.py:45
REP_NUM_CORRECTION = 20    # tracker.py:46
BOUNDARY_XY = 6            # tracker.py:47
K_POINTS = 20              # tracker.py:1259


def read_image_ts(vol, path, name, z_range, print_=False):
    """tracker.py:113-142: stack z slices <path>/<name % (vol, z)> into (x, y, z)."""
    from PIL import Image
    image_raw = []
    for z in range(z_range[0], z_range[1]):
        image_raw.append(np.array(Image.open(os.path.join(path, name % (vol, z)))))
    img_array = np.array(image_raw).transpose((1, 2, 0))
    if print_:
        print("Load images with shape:", img_array.shape)
    return img_array


class History:
    """tracker.py:756-776."""

    def __init__(self):
        self.r_displacements = []
        self.r_segmented_coordinates = []
        self.r_tracked_coordinates = []
        self.anim = []


class SegResults:
    """tracker.py:463-497."""

    def __init__(self):
        self.image_cell_bg = None
        self.l_center_coordinates = None
        self.segmentation_auto = None
        self.image_gcn = None
        self.r_coordinates_segment = None

    def update_results(self, image_cell_bg, l_center_coordinates, segmentation_auto, image_gcn, r_coordinates_segment):
        self.image_cell_bg = image_cell_bg
        self.l_center_coordinates = l_center_coordinates
        self.segmentation_auto = segmentation_auto
        self.image_gcn = image_gcn
        self.r_coordinates_segment = r_coordinates_segment


class Tracker:
    """Constructor arguments follow tracker.py:854-859.  `image_source` (extra, optional) is a callable
    vol -> (x, y, z) ndarray that replaces TIFF reading for in-memory / synthetic stacks."""

    def __init__(self, volume_num, siz_xyz, z_xy_ratio, z_scaling, noise_level, min_size, beta_tk, lambda_tk,
                 maxiter_tk, folder_path=None, image_name=None, unet_model_file=None, ffn_model_file=None,
                 cell_num=0, ensemble=False, adjacent=False, shrink=(24, 24, 2), miss_frame=None, image_source=None):
        self.volume_num = volume_num
        self.x_siz, self.y_siz, self.z_siz = siz_xyz
        self.z_xy_ratio = z_xy_ratio
        self.z_scaling = z_scaling
        self.noise_level = noise_level
        self.min_size = min_size
        self.beta_tk = beta_tk
        self.lambda_tk = lambda_tk
        self.max_iteration = maxiter_tk
        self.folder_path = folder_path
        self.image_name = image_name
        self.unet_model_file = unet_model_file
        self.ffn_model_file = ffn_model_file
        self.cell_num = cell_num
        self.ensemble = ensemble
        self.adjacent = adjacent
        self.shrink = shrink
        self.miss_frame = [] if not miss_frame else miss_frame
        self.image_source = image_source
        self.use_8_bit = True if cell_num <= 255 else False                      # tracker.py:881
        # folder layout + float16 U-Net cache of the reference (tracker.py:652-669,734-752) when a folder is given
        self.paths = io_formats.Paths(folder_path, image_name, unet_model_file, ffn_model_file)
        self.cache_unet_regions = folder_path is not None
        if folder_path is not None:
            self.paths.make_folders(adjacent, ensemble)
        self.unet_model = None
        self.ffn_model = None
        self.vol = None
        self.segresult = SegResults()
        self.history = History()
        self.cell_num_t0 = None
        self.r_coordinates_tracked_t0 = None
        self.r_coordinates_segment_t0 = None
        self.cells_on_boundary = None
        self.keep_on_device = False          # True: _segment leaves probability map and label image in HBM
        self._last_seg_device = None
        # accurate correction (tracker.py:1044-1110): state prepared by interpolate_seg + cal_subregions
        self.segmentation_manual_relabels = None
        self.seg_cells_interpolated_corrected = None
        self.Z_RANGE_INTERP = None
        self.region_cells = None
        self.tracked_labels = None
        self._unet_cache = {}
        self._seg_stream = None

    # ------------------------------------------------------------------ models
    def load_unet(self, model=None):
        """tracker.py:575-581.  Accepts a UNet3 instance or loads <folder>/models/<unet_model_file> (npz)."""
        if model is not None:
            self.unet_model = model
        else:
            self.unet_model = UNet3("a")
            self.unet_model.load_weights(os.path.join(self.folder_path, "models", self.unet_model_file))

    def load_ffn(self, model=None):
        """tracker.py:1119-1122."""
        if model is not None:
            self.ffn_model = model
        else:
            self.ffn_model = FFN()
            self.ffn_model.load_weights(os.path.join(self.folder_path, "models", self.ffn_model_file))

    def set_segmentation(self, noise_level=None, min_size=None, del_cache=False):
        """tracker.py:520-550 (cache = in-memory dict here)."""
        changed = False
        if noise_level is not None and noise_level != self.noise_level:
            self.noise_level, changed = noise_level, True
        if min_size is not None and min_size != self.min_size:
            self.min_size, changed = min_size, True
        if changed or del_cache:
            self._unet_cache.clear()
            if self.cache_unet_regions and self.paths.unet_cache:
                for f in os.listdir(self.paths.unet_cache):                     # tracker.py:543-547
                    os.remove(os.path.join(self.paths.unet_cache, f))

    def set_tracking(self, beta_tk, lambda_tk, maxiter_tk):
        """tracker.py:889-917."""
        self.beta_tk, self.lambda_tk, self.max_iteration = beta_tk, lambda_tk, maxiter_tk

    # ------------------------------------------------------------------ segmentation
    def _read_raw(self, vol):
        if self.image_source is not None:
            return np.asarray(self.image_source(vol))
        return read_image_ts(vol, os.path.join(self.folder_path, "data"), self.image_name, (1, self.z_siz + 1))

    def _transform_layer_to_real(self, voxel_disp):
        new_disp = np.array(voxel_disp, dtype=np.float64).copy()
        new_disp[:, 2] = new_disp[:, 2] * self.z_xy_ratio
        return new_disp

    def _save_unet_regions(self, image_raw, vol):
        """tracker.py:662-669: LCN normalisation + tiled U-Net; volume stays in HBM between the two."""
        raw_dev = _raw_to_device(image_raw)
        norm_dev = normalize_image_device(raw_dev, self.noise_level, (27, 27, 1))
        prob_dev = self.unet_model.prediction_device(norm_dev, self.shrink)
        return prob_dev

    def _enqueue_segmentation(self, image_raw, vol):
        """LCN + U-Net of `vol` on the segmentation stream (all segmentation work is serialised there: it shares
        one set of workspaces); returns immediately."""
        if self._seg_stream is None:
            self._seg_stream = torch.cuda.Stream()
        self._seg_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._seg_stream):
            prob_dev = self._save_unet_regions(image_raw, vol)
            ready = torch.cuda.Event()
            ready.record(self._seg_stream)
        self._unet_cache[vol] = (prob_dev, ready)

    def _predict_cellregions_device(self, image_raw, vol):
        """tracker.py:652-660 with the result left in HBM (first-pass fp32 result; see `cache_unet_regions` for the
        reference's fp16 disk cache).  A result prefetched by `prefetch_segmentation` is picked up here."""
        if vol not in self._unet_cache and self.cache_unet_regions:
            cached = io_formats.load_unet_cache(self.paths.unet_cache, vol)     # float16 (1, x, y, z, 1), tracker.py:656
            if cached is not None:
                return to_device(np.ascontiguousarray(cached[0, :, :, :, 0]).astype(np.float32), torch.float32)
        fresh = vol not in self._unet_cache
        if fresh:
            self._enqueue_segmentation(image_raw, vol)
        prob_dev, ready = self._unet_cache[vol]
        torch.cuda.current_stream().wait_event(ready)
        self._unet_cache = {v: e for v, e in self._unet_cache.items() if v >= vol}
        if self.cache_unet_regions:
            io_formats.save_unet_cache(self.paths.unet_cache, vol, prob_dev.cpu().numpy()[None, ..., None])
        return prob_dev

    def _predict_cellregions(self, image_raw, vol):
        """tracker.py:652-660: host array (1, x, y, z, 1)."""
        return self._predict_cellregions_device(image_raw, vol).cpu().numpy()[None, ..., None]

    def prefetch_segmentation(self, vol):
        """Enqueue LCN + U-Net of volume `vol` so that it overlaps the match + track stage of the volume before it
        (pipeline.py explains why the two stages pair well).  Used by `track`."""
        if vol > self.volume_num or vol in self.miss_frame or vol in self._unet_cache:
            return
        self._enqueue_segmentation(self._read_raw(vol), vol)

    def _watershed_device(self, prob_dev, method):
        """tracker.py:671-684 on the device: watershed_2d + watershed_3d + relabel_sequential and the centres of mass
        (tracker.py:646-648) in one call; updates min_size / cell_num like the reference."""
        seg = _ws.segment_device(prob_dev, self.z_xy_ratio, method, self.min_size, self.cell_num)
        n, min_size, cell_num = seg.host_scalars()
        self.min_size = min_size
        if method == "min_size":
            self.cell_num = cell_num
        return seg, n

    def _watershed(self, image_cell_bg, method):
        """tracker.py:671-684, host-array form: (1, x, y, z, 1) probabilities -> segmentation_auto (x, y, z)."""
        prob_dev = torch.from_numpy(np.ascontiguousarray(image_cell_bg[0, :, :, :, 0], dtype=np.float32)).cuda()
        seg, _ = self._watershed_device(prob_dev, method)
        return seg.labels.cpu().numpy()

    def _segment(self, vol, method="min_size", print_shape=False):
        """tracker.py:605-650.  The probability map, the label image and the centres are produced on the device; the
        host copies the reference returns (image_cell_bg, segmentation_auto) are made here because callers of the
        class API read them (set `keep_on_device` to skip the two big downloads inside `track`)."""
        image_raw = self._read_raw(vol)
        self._last_raw = image_raw
        image_gcn = image_raw.copy() / 65536.0
        prob_dev = self._predict_cellregions_device(image_raw, vol)
        seg, n = self._watershed_device(prob_dev, method)
        if n == 0:
            if float(prob_dev.max()) <= 0.5:
                raise ValueError("No cell was detected by 3D U-Net! Try to reduce the noise_level.")
            raise ValueError("No cell was detected by watershed! Try to reduce the min_size.")
        l_center_coordinates = [tuple(c) for c in seg.centres_host()]
        r_coordinates_segment = self._transform_layer_to_real(l_center_coordinates)
        self._last_seg_device = (prob_dev, seg.labels)
        if self.keep_on_device:
            image_cell_bg, segmentation_auto = None, None
        else:
            image_cell_bg = prob_dev.cpu().numpy()[None, ..., None]
            segmentation_auto = seg.labels.cpu().numpy()
        return image_cell_bg, l_center_coordinates, segmentation_auto, image_gcn, r_coordinates_segment

    def segment_vol1(self, method="min_size"):
        """tracker.py:583-603."""
        self.vol = 1
        self.segresult.update_results(*self._segment(self.vol, method=method, print_shape=True))
        self.r_coordinates_segment_t0 = self.segresult.r_coordinates_segment.copy()

    def initiate_tracking(self, r_coordinates_tracked_t0=None):
        """tracker.py:1124-1136 reduced to its state initialisation: volume 1's (proof-read) cell centres become
        the tracked set; without manual correction the automatic segmentation is used."""
        if r_coordinates_tracked_t0 is None:
            r_coordinates_tracked_t0 = self.r_coordinates_segment_t0
        self.r_coordinates_tracked_t0 = np.asarray(r_coordinates_tracked_t0, dtype=np.float64).copy()
        self.cell_num_t0 = self.r_coordinates_tracked_t0.shape[0]
        self.cells_on_boundary = np.zeros(self.cell_num_t0, dtype=int)
        self.history = History()
        self.history.r_displacements.append(np.zeros((self.cell_num_t0, 3)))
        self.history.r_segmented_coordinates.append(self.r_coordinates_segment_t0)
        self.history.r_tracked_coordinates.append(self.r_coordinates_tracked_t0)

    # ------------------------------------------------------------------ accurate correction (volume-1 preparation)
    def load_manual_seg(self, segmentation=None):
        """tracker.py:934-950: the proof-read segmentation of volume 1, from `manual_vol1/*.tif` or given directly
        (e.g. `segresult.segmentation_auto` when no manual correction is made); relabelled sequentially."""
        if segmentation is None:
            files = sorted(f for f in os.listdir(self.paths.manual_segmentation_vol1) if f.lower().endswith((".tif", ".tiff")))
            from PIL import Image
            segmentation = np.array([np.array(Image.open(os.path.join(self.paths.manual_segmentation_vol1, f)))
                                     for f in files]).transpose((1, 2, 0))
        seg = np.asarray(segmentation)
        uniq = np.unique(seg)
        uniq = uniq[uniq != 0]
        fw = np.zeros(int(seg.max()) + 1, dtype=np.int64)
        fw[uniq] = np.arange(1, len(uniq) + 1)
        self.segmentation_manual_relabels = fw[seg]
        if self.segmentation_manual_relabels.max() > 255:
            self.use_8_bit = False

    def _interpolate(self):
        """tracker.py:1083-1091."""
        seg_interp, seg_cover = _corr.interpolate_labels(self.segmentation_manual_relabels, z_scaling=self.z_scaling,
                                                         smooth_sigma=2.5)
        corrected = _corr.recalculate_cell_boundaries(seg_interp, seg_cover)
        return corrected[5:self.x_siz + 5, 5:self.y_siz + 5, 5:self.z_siz * self.z_scaling + 5]

    def interpolate_seg(self):
        """tracker.py:1044-1071: interpolate / smooth the cells of volume 1 along z, re-draw their boundaries, split
        regions that fell apart, and take the cell centres as the tracked set."""
        self.seg_cells_interpolated_corrected = self._interpolate()
        self.Z_RANGE_INTERP = range(self.z_scaling // 2, self.seg_cells_interpolated_corrected.shape[2], self.z_scaling)
        num_cells = np.size(np.unique(self.seg_cells_interpolated_corrected)) - 1
        relabelled, found = _corr.label_components(self.seg_cells_interpolated_corrected)
        if num_cells != found:
            print(f"WARNING: {num_cells} cells were manually labeled while the program found {found} separated cells "
                  f"and corrected it")
        self.seg_cells_interpolated_corrected = relabelled.astype(np.int64)
        self.segmentation_manual_relabels = self.seg_cells_interpolated_corrected[:, :, self.Z_RANGE_INTERP]
        if self.folder_path is not None:
            io_formats.save_img3ts(range(0, self.z_siz), self.segmentation_manual_relabels,
                                   self.paths.track_results + "track_results_t%06i_z%04i.tif", t=1, use_8_bit=self.use_8_bit)
        seg = _ws.segment_centres(self.segmentation_manual_relabels)
        self.r_coordinates_tracked_t0 = self._transform_layer_to_real(seg)
        self.cell_num_t0 = self.r_coordinates_tracked_t0.shape[0]

    def cal_subregions(self):
        """tracker.py:1093-1110: per-cell voxel lists for the quick accurate correction (device resident)."""
        self.region_cells = _corr.CellRegions(self.seg_cells_interpolated_corrected, self.z_scaling)
        self.region_xyz_min, self.region_width = self.region_cells.region_xyz_min, self.region_cells.region_width
        self.pad_x, self.pad_y, self.pad_z = (int(v) for v in self.region_cells.pad)

    def _transform_real_to_interpolated(self, r_disp):
        new_disp = np.array(r_disp, dtype=np.float64).copy()                  # tracker.py:563-565
        new_disp[:, 2] = new_disp[:, 2] * (self.z_scaling / self.z_xy_ratio)
        return np.rint(new_disp).astype(int)

    def _accurate_correction(self, cells_on_boundary_local, r_coor_predicted):
        """tracker.py:1177-1191 (the whole repetition loop is one device call).  Needs `_segment` of the target volume
        (probability map + raw stack on the device) and `cal_subregions`."""
        prob_dev, _ = self._last_seg_device
        raw_dev = _raw_to_device(self._last_raw)
        r_disp, i_disp, _ = _corr.accurate_correction_device(
            self.region_cells, prob_dev, raw_dev, self.z_xy_ratio, self.r_coordinates_tracked_t0,
            self.history.r_displacements[-1], self.history.r_tracked_coordinates[-1], r_coor_predicted,
            cells_on_boundary_local, REP_NUM_CORRECTION)
        return r_disp.cpu().numpy(), i_disp.cpu().numpy().astype(int)

    def _transform_motion_to_image(self, cells_on_boundary_local, i_disp_from_vol1_updated):
        """tracker.py:1391-1399."""
        lab = _corr.tracked_labels_device(self.region_cells, i_disp_from_vol1_updated, cells_on_boundary_local,
                                          (self.x_siz, self.y_siz, self.z_siz))
        return lab.cpu().numpy().astype(np.int64)

    # ------------------------------------------------------------------ FFN + PR-GLS
    def _fit_predict_batch(self, source_vols):
        """_fit_ffn_prgls + _predict_one_rep (tracker.py:1193-1289) for ALL source volumes at once.

        Per repetition i: one FFN match per member, then ONE batched EM launch (one CTA per member) with
        beta * 0.8**i; the fitted transforms are replayed on the tracked cells.  Everything stays on the
        device; returns a (E, L, 3) float64 CUDA tensor of predicted coordinates."""
        tgt_dev = to_device(np.asarray(self.segresult.r_coordinates_segment, dtype=np.float64), torch.float64)
        inter = [to_device(np.asarray(self.history.r_segmented_coordinates[v - 1], dtype=np.float64), torch.float64)
                 for v in source_vols]
        pred = [to_device(np.asarray(self.history.r_tracked_coordinates[v - 1], dtype=np.float64), torch.float64)
                for v in source_vols]
        self._last_fit = []
        for i in range(REP_NUM_PRGLS):
            beta = self.beta_tk * (0.8 ** i)
            probs = []
            for e in range(len(source_vols)):
                corr = self.ffn_model.match_device(inter[e], tgt_dev, K_POINTS)      # initial_matching_quick
                probs.append(EmProblem(inter[e], tgt_dev, corr))
            run_em(probs, MODE_TRACK, beta, self.lambda_tk, self.max_iteration, 1e8, 0.5)   # pr_gls_quick
            for e, p in enumerate(probs):
                pred[e] = predict_one_rep_device(pred[e], inter[e], beta, p.coef)   # _predict_one_rep
                inter[e] = p.ref_out                                               # next repetition starts at T_X
            self._last_fit.append(probs)
        return torch.stack(pred, dim=0)

    def _fit_ffn_prgls(self, rep, r_coordinates_segment_pre):
        """tracker.py:1224-1254 (host-array form): returns (C_t, BETA_t, coor_intermediate_list)."""
        tgt_dev = to_device(np.asarray(self.segresult.r_coordinates_segment, dtype=np.float64), torch.float64)
        inter = to_device(np.asarray(r_coordinates_segment_pre, dtype=np.float64), torch.float64)
        C_t, BETA_t, coor_list = [], [], []
        for i in range(rep):
            beta = self.beta_tk * (0.8 ** i)
            coor_list.append(inter.cpu().numpy())
            corr = self.ffn_model.match_device(inter, tgt_dev, K_POINTS)
            p = run_em([EmProblem(inter, tgt_dev, corr)], MODE_TRACK, beta, self.lambda_tk, self.max_iteration,
                       1e8, 0.5)[0]
            C_t.append(p.coef.cpu().numpy())
            BETA_t.append(beta)
            inter = p.ref_out
        return C_t, BETA_t, coor_list

    def _predict_one_rep(self, r_coordinates_predicted_pre, coor_intermediate_list, BETA_t, C_t):
        """tracker.py:1269-1289."""
        pre = to_device(np.asarray(r_coordinates_predicted_pre, dtype=np.float64), torch.float64)
        inter = to_device(np.asarray(coor_intermediate_list, dtype=np.float64), torch.float64)
        coef = to_device(np.asarray(C_t, dtype=np.float64), torch.float64)
        post = predict_one_rep_device(pre, inter, BETA_t, coef)
        return post.cpu().numpy(), np.asarray(r_coordinates_predicted_pre)

    def _predict_pos_once(self, source_volume, draw=False):
        """tracker.py:1193-1222 (no animation)."""
        return self._fit_predict_batch([source_volume])[0].cpu().numpy(), None

    def _get_cells_onBoundary(self, r_coordinates_prgls, ensemble):
        """tracker.py:1291-1308."""
        boundary_xy = 0 if ensemble else BOUNDARY_XY
        c = r_coordinates_prgls
        return np.where((c[:, 0] < boundary_xy) | (c[:, 1] < boundary_xy) |
                        (c[:, 0] > self.x_siz - boundary_xy) | (c[:, 1] > self.y_siz - boundary_xy) |
                        (c[:, 2] / self.z_xy_ratio < 0) | (c[:, 2] / self.z_xy_ratio > self.z_siz))[0]

    # ------------------------------------------------------------------ public loop
    def match(self, target_volume, method="min_size"):
        """tracker.py:1138-1175 without drawing / accurate correction."""
        if target_volume in self.miss_frame:
            raise ValueError("target_volume is a miss_frame")
        self.segresult.update_results(*self._segment(target_volume, method=method))
        r_coor_predicted, anim = self._predict_pos_once(source_volume=1, draw=True)
        cells_bd = self._get_cells_onBoundary(r_coor_predicted, self.ensemble)
        cells_on_boundary_local = self.cells_on_boundary.copy()
        cells_on_boundary_local[cells_bd] = 1
        return anim, [cells_on_boundary_local, target_volume, None, r_coor_predicted]

    def track_one_vol(self, target_volume, fig=None, axc6=None, method="min_size"):
        """tracker.py:1473-1536 (prediction + history update; label-image output is host glue, omitted)."""
        if target_volume in self.miss_frame:
            self.history.r_displacements.append(self.history.r_displacements[-1])
            self.history.r_segmented_coordinates.append(self.segresult.r_coordinates_segment)
            self.history.r_tracked_coordinates.append(self.r_coordinates_tracked_t0 + self.history.r_displacements[-1])
            return None
        self.segresult.update_results(*self._segment(target_volume, method=method))
        source_vols_list = get_reference_vols(self.ensemble, target_volume, adjacent=self.adjacent)
        stack = self._fit_predict_batch(source_vols_list)
        r_coor_predicted_mean = trim_mean_device(stack, 0.1).cpu().numpy()
        cells_bd = self._get_cells_onBoundary(r_coor_predicted_mean, self.ensemble)
        self.cells_on_boundary[cells_bd] = 1
        if self.region_cells is not None:
            # accurate correction + tracked label image (tracker.py:1512-1522), on the device
            r_disp_from_vol1_updated, i_disp_from_vol1_updated = \
                self._accurate_correction(self.cells_on_boundary, r_coor_predicted_mean)
            self.tracked_labels = self._transform_motion_to_image(self.cells_on_boundary, i_disp_from_vol1_updated)
            if self.folder_path is not None:
                io_formats.save_img3ts(range(0, self.z_siz), self.tracked_labels,
                                       self.paths.track_results + "track_results_t%06i_z%04i.tif", target_volume, self.use_8_bit)
        else:
            # without interpolate_seg + cal_subregions the positions stay those of FFN + PR-GLS
            r_disp_from_vol1_updated = self.history.r_displacements[-1] + \
                (r_coor_predicted_mean - self.history.r_tracked_coordinates[-1])
        if self.ensemble:
            self.cells_on_boundary = np.zeros(self.cell_num_t0).astype(int)
        self.history.r_displacements.append(r_disp_from_vol1_updated)
        self.history.r_segmented_coordinates.append(self.segresult.r_coordinates_segment)
        self.history.r_tracked_coordinates.append(self.r_coordinates_tracked_t0 + r_disp_from_vol1_updated)
        return None

    def _reset_tracking_state(self, from_volume):
        """tracker.py:1462-1471."""
        assert from_volume >= 2, "from_volume should >= 2"
        current_vol = len(self.history.r_displacements)
        del self.history.r_displacements[from_volume - 1:]
        del self.history.r_segmented_coordinates[from_volume - 1:]
        del self.history.r_tracked_coordinates[from_volume - 1:]
        assert len(self.history.r_displacements) == from_volume - 1, \
            f"Currently data has been tracked until vol {current_vol}, the program cannot start from {from_volume}"

    def track(self, fig=None, ax=None, from_volume=2):
        """tracker.py:1415-1431."""
        self._reset_tracking_state(from_volume)
        lib = _lib.lib()
        old = lib.ct_set_reserved_sms(1)          # the EM's single CTA keeps an SM while the next volume is segmented
        try:
            for vol in range(from_volume, self.volume_num + 1):
                self.prefetch_segmentation(vol + 1)
                self.track_one_vol(vol, fig, ax)
        finally:
            lib.ct_set_reserved_sms(old)
        return None

    def save_coordinates(self, path=None):
        """tracker.py:1538-1551: CSV of tracked coordinates (t, cell, x, y, z)."""
        path = path or os.path.join(self.paths.track_information, "tracked_coordinates.csv")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        coord = np.asarray(self.history.r_tracked_coordinates)
        t, cell = np.meshgrid(np.arange(1, coord.shape[0] + 1), np.arange(1, coord.shape[1] + 1), indexing="ij")
        table = np.column_stack([t.ravel(), cell.ravel(), coord.reshape(-1, 3)])
        np.savetxt(path, table, delimiter=",", header="t,cell,x,y,z", comments="")
