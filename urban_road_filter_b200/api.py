"""ctypes binding of liburf_b200.so plus `Detector`, the host-side mirror of the reference node's interface
(`paramsCallback` -> set_params, `Detector::filtered` -> filtered / filtered_batch; src/main.cpp:4-34,
src/lidar_segmentation.cpp:95). There is no CPU fallback: loading fails loudly when the CUDA library is missing and
every compute call raises without a GPU."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .ctypes_abi import (UrfMqStats, QUEUE_PROCESS_FN, URF_ERR_CLOSED, URF_ERR_TIMEOUT, URF_MAX_CHANNELS, URF_MAX_VERTS, URF_OK, URF_QUEUE_BLOCK,
                         URF_QUEUE_DROP_OLDEST, URF_TOO_FEW_POINTS, UrfClouds, UrfParams, UrfQueueStats, UrfResult, UrfStrip,
                         make_params)

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "liburf_b200.so")

EXPORTS = ["urf_queue_next_view", "urf_queue_release_view", "urf_mq_next_view", "urf_queue_submit_ref", "urf_queue_create_cloud2", "urf_queue_submit_cloud2", "urf_mq_create", "urf_mq_create_with",
           "urf_mq_set_params", "urf_mq_submit", "urf_mq_submit_ref", "urf_mq_next", "urf_mq_get_stats", "urf_mq_close", "urf_mq_destroy",
           "urf_process_cloud2", "urf_process_cloud2_packed", "urf_pinned_alloc", "urf_pinned_free", "urf_queue_create",
           "urf_queue_create_with", "urf_queue_submit", "urf_queue_next", "urf_queue_get_stats", "urf_queue_close", "urf_queue_destroy",
           "urf_version", "urf_strerror", "urf_last_cuda_error", "urf_default_params", "urf_create", "urf_destroy",
           "urf_set_params", "urf_get_params", "urf_process", "urf_process_batch", "urf_process_batch_device",
           "urf_process_batch_xyz", "urf_process_cloud2_batch", "urf_enqueue_batch_device", "urf_enqueue_batch_device_ex", "urf_finish_batch_device", "urf_stream", "urf_last_device_ms",
           "urf_last_launch_count", "urf_build_markers"]

_lib = None


class UrfError(RuntimeError):
    def __init__(self, code: int, where: str, detail: str = ""):
        self.code = code
        super().__init__(f"{where}: urf error {code}" + (f" ({detail})" if detail else ""))


def load_library(path: str = LIB_PATH) -> C.CDLL:
    """Loads the CUDA library. Raises if it has not been built — there is deliberately no fallback."""
    global _lib
    if _lib is not None and path == LIB_PATH:
        return _lib
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(urban_road_filter_b200 has no CPU fallback)")
    lib = C.CDLL(path)
    vp, ip = C.c_void_p, C.c_int
    lib.urf_version.restype = ip
    lib.urf_strerror.restype = C.c_char_p
    lib.urf_strerror.argtypes = [ip]
    lib.urf_last_cuda_error.restype = C.c_char_p
    lib.urf_last_cuda_error.argtypes = [vp]
    lib.urf_default_params.argtypes = [C.POINTER(UrfParams)]
    lib.urf_create.argtypes = [C.POINTER(vp), ip, ip, ip]
    lib.urf_destroy.argtypes = [vp]
    lib.urf_set_params.argtypes = [vp, C.POINTER(UrfParams)]
    lib.urf_get_params.argtypes = [vp, C.POINTER(UrfParams)]
    lib.urf_set_option.argtypes = [vp, ip, ip]
    lib.urf_process.argtypes = [vp, vp, ip, C.POINTER(UrfResult)]
    lib.urf_process_batch.argtypes = [vp, C.POINTER(vp), C.POINTER(ip), ip, C.POINTER(UrfResult)]
    lib.urf_process_cloud2.argtypes = [vp, vp, ip, ip, ip, ip, ip, C.POINTER(UrfResult)]
    lib.urf_process_cloud2_packed.argtypes = [vp, vp, ip, ip, ip, ip, ip, ip, C.POINTER(UrfResult), C.POINTER(UrfClouds)]
    lib.urf_process_batch_xyz.argtypes = [vp, C.POINTER(vp), C.POINTER(ip), ip, C.POINTER(UrfResult), C.POINTER(vp)]
    lib.urf_process_cloud2_batch.argtypes = [vp, C.POINTER(vp), C.POINTER(ip), ip, ip, ip, ip, ip, ip, C.POINTER(UrfResult), C.POINTER(vp)]
    lib.urf_process_batch_device.argtypes = [vp, vp, ip, C.POINTER(ip), ip, vp, C.POINTER(UrfResult)]
    lib.urf_enqueue_batch_device.argtypes = [vp, vp, ip, C.POINTER(ip), ip, vp]
    lib.urf_enqueue_batch_device_ex.argtypes = [vp, vp, ip, C.POINTER(ip), ip, vp, vp]
    lib.urf_finish_batch_device.argtypes = [vp, C.POINTER(UrfResult)]
    lib.urf_stream.restype = vp
    lib.urf_stream.argtypes = [vp]
    lib.urf_last_device_ms.restype = C.c_float
    lib.urf_last_device_ms.argtypes = [vp]
    lib.urf_last_launch_count.argtypes = [vp]
    lib.urf_build_markers.argtypes = [C.POINTER(UrfParams), vp, ip, C.POINTER(ip), C.POINTER(UrfStrip), ip, vp, ip,
                                      C.POINTER(ip)]
    lib.urf_pinned_alloc.restype = vp
    lib.urf_pinned_alloc.argtypes = [C.c_size_t]
    lib.urf_pinned_free.restype = None
    lib.urf_pinned_free.argtypes = [vp]
    lib.urf_queue_create.argtypes = [C.POINTER(vp), vp, ip, ip, ip, ip]
    lib.urf_queue_create_with.argtypes = [C.POINTER(vp), QUEUE_PROCESS_FN, vp, ip, ip, ip, ip]
    lib.urf_queue_submit.argtypes = [vp, vp, ip, C.c_uint64, ip]
    lib.urf_queue_next.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(UrfResult), ip]
    lib.urf_queue_get_stats.argtypes = [vp, C.POINTER(UrfQueueStats)]
    lib.urf_queue_close.restype = None
    lib.urf_queue_close.argtypes = [vp]
    lib.urf_queue_destroy.restype = None
    lib.urf_queue_destroy.argtypes = [vp]
    lib.urf_queue_submit_ref.argtypes = [vp, vp, ip, C.c_uint64, ip]
    lib.urf_queue_create_cloud2.argtypes = [C.POINTER(vp), vp, ip, ip, ip, ip, ip, ip, ip, ip, ip]
    lib.urf_queue_submit_cloud2.argtypes = [vp, vp, ip, C.c_uint64, ip]
    lib.urf_mq_create.argtypes = [C.POINTER(vp), C.POINTER(ip), ip, ip, ip, ip, C.POINTER(UrfParams)]
    lib.urf_mq_create_with.argtypes = [C.POINTER(vp), QUEUE_PROCESS_FN, C.POINTER(vp), ip, ip, ip, ip]
    lib.urf_mq_set_params.argtypes = [vp, C.POINTER(UrfParams)]
    lib.urf_mq_submit.argtypes = [vp, vp, ip, C.c_uint64, ip]
    lib.urf_mq_submit_ref.argtypes = [vp, vp, ip, C.c_uint64, ip]
    lib.urf_mq_next.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(UrfResult), ip]
    lib.urf_mq_get_stats.argtypes = [vp, C.POINTER(UrfMqStats)]
    lib.urf_mq_close.argtypes = [vp]
    lib.urf_mq_close.restype = None
    lib.urf_mq_destroy.argtypes = [vp]
    lib.urf_mq_destroy.restype = None
    lib.urf_test_math.argtypes = [ip, ip, vp, vp, vp, ip]
    lib.urf_debug_fetch.argtypes = [vp, ip, ip, vp, C.c_size_t]
    lib.urf_debug_sizeof_tab.restype = C.c_size_t
    lib.urf_profile_count.argtypes = [vp]
    lib.urf_profile_slots.argtypes = [vp]
    lib.urf_profile_get.argtypes = [vp, ip, ip, C.POINTER(C.c_char_p), C.POINTER(C.c_float)]
    if path == LIB_PATH:
        _lib = lib
    return lib


class ScanResult:
    """Per-scan output of the path (urf_result, include/urf.h)."""
    __slots__ = ("status", "n_in", "n_roi", "n_rings", "n_order", "n_road", "n_curb", "n_vert", "flags", "label",
                 "ring", "order", "ring_start", "vert")

    @property
    def published(self) -> bool:
        return self.status == URF_OK

    def cloud_indices(self, which: str) -> np.ndarray:
        """Input indices of the `road` / `curb` / `road_probably` / `roi` clouds in the reference's emission order
        (lidar_segmentation.cpp:354-367,605-608,620)."""
        if which == "roi":
            return np.nonzero(self.label >= 0)[0].astype(np.int32)
        if self.order is None:
            raise ValueError("emission order was not requested")
        if which == "road_probably":
            if self.n_rings <= 10:
                return np.zeros(0, np.int32)
            return self.order[self.ring_start[10]: self.ring_start[11]]
        lab = self.label[self.order]
        return self.order[lab == (1 if which == "road" else 2)]


def build_markers(prm: UrfParams, vert: np.ndarray, ghostcount: int = 0):
    """Marker tail (lidar_segmentation.cpp:371-598). Returns (strips, ghostcount'): strips = [(id, action, red, xyz[n,3])]."""
    lib = load_library()
    vert = np.ascontiguousarray(vert, np.float32).reshape(-1, 4)
    strips = (UrfStrip * 1024)()
    pts = np.zeros(3 * 4096, np.float64)
    gc = C.c_int(ghostcount)
    npnt = C.c_int(0)
    ns = lib.urf_build_markers(C.byref(prm), vert.ctypes.data, vert.shape[0], C.byref(gc), strips, 1024,
                               pts.ctypes.data, 4096, C.byref(npnt))
    if ns < 0:
        raise UrfError(ns, "urf_build_markers")
    out = [(s.id, s.action, s.red, pts[3 * s.first: 3 * (s.first + s.count)].reshape(-1, 3).copy()) for s in strips[:ns]]
    return out, gc.value


class Detector:
    """Host-side mirror of the reference's `Detector` (include/urban_road_filter/data_structures.hpp:110-141) on one GPU."""

    def __init__(self, max_points: int, max_batch: int = 1, device: int = 0, params: UrfParams | None = None):
        self.lib = load_library()
        self._ctx = C.c_void_p()
        rc = self.lib.urf_create(C.byref(self._ctx), device, max_points, max_batch)
        if rc != URF_OK:
            raise UrfError(rc, "urf_create", self.lib.urf_strerror(rc).decode())
        self.max_points, self.max_batch, self.device = max_points, max_batch, device
        self.params = params if params is not None else make_params()
        self.set_params(self.params)
        self.ghostcount = 0      # lidar_segmentation.cpp:23

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self.lib.urf_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, where: str):
        if rc < 0:
            raise UrfError(rc, where, self.lib.urf_strerror(rc).decode() + ": " + self.lib.urf_last_cuda_error(self._ctx).decode())

    def set_params(self, prm: UrfParams):
        """paramsCallback (src/main.cpp:4-34)."""
        self._check(self.lib.urf_set_params(self._ctx, C.byref(prm)), "urf_set_params")
        self.params = prm

    def set_option(self, option: int, value: int):
        self._check(self.lib.urf_set_option(self._ctx, option, value), "urf_set_option")

    def filtered_batch(self, clouds, want_ring: bool = True, want_order: bool = True) -> list[ScanResult]:
        """`batch` independent Detector::filtered() calls (lidar_segmentation.cpp:95) on host (N,4) float32 arrays."""
        B = len(clouds)
        arrs = [np.ascontiguousarray(c, np.float32).reshape(-1, 4) for c in clouds]
        ptrs = (C.c_void_p * B)(*[a.ctypes.data for a in arrs])
        ns = (C.c_int * B)(*[a.shape[0] for a in arrs])
        res = (UrfResult * B)()
        keep = []
        for b, a in enumerate(arrs):
            m = max(a.shape[0], 1)
            lab = np.full(m, -1, np.int32)
            ring = np.full(m, -1, np.int32) if want_ring else None
            order = np.zeros(m, np.int32) if want_order else None
            rs = np.zeros(URF_MAX_CHANNELS + 1, np.int32)
            res[b].label = lab.ctypes.data_as(C.POINTER(C.c_int32))
            if want_ring:
                res[b].ring = ring.ctypes.data_as(C.POINTER(C.c_int32))
            if want_order:
                res[b].order = order.ctypes.data_as(C.POINTER(C.c_int32))
            res[b].ring_start = rs.ctypes.data_as(C.POINTER(C.c_int32))
            keep.append((lab, ring, order, rs))
        self._check(self.lib.urf_process_batch(self._ctx, ptrs, ns, B, res), "urf_process_batch")
        out = []
        for b, a in enumerate(arrs):
            lab, ring, order, rs = keep[b]
            r = ScanResult()
            n = a.shape[0]
            for f in ("status", "n_in", "n_roi", "n_rings", "n_order", "n_road", "n_curb", "n_vert", "flags"):
                setattr(r, f, int(getattr(res[b], f)))
            r.label = lab[:n]
            r.ring = ring[:n] if want_ring else None
            r.order = order[: r.n_order].copy() if want_order else None
            r.ring_start = rs[: r.n_rings + 1].copy()
            r.vert = np.ctypeslib.as_array(res[b].vert).reshape(URF_MAX_VERTS, 4)[: r.n_vert].copy()
            out.append(r)
        return out

    def filtered_batch_records(self, records, point_step: int, off_x: int, off_y: int, off_z: int, off_intensity: int = -1,
                               want_order: bool = False, label8: bool = True) -> list[ScanResult]:
        """`batch` scans given as raw PointCloud2 record arrays (uint8, n * point_step bytes each) of one sensor format,
        unpacked on the device (urf_process_cloud2_batch). point_step == 12 with offsets 0, 4, 8 is the packed-xyz lean
        input of urf_process_batch_xyz, which this method then calls. label8: labels come back as int8."""
        B = len(records)
        raws = [np.ascontiguousarray(r).view(np.uint8).reshape(-1) for r in records]
        ns = [r.size // point_step for r in raws]
        ptrs = (C.c_void_p * B)(*[r.ctypes.data for r in raws])
        cn = (C.c_int * B)(*ns)
        res = (UrfResult * B)()
        keep = []
        l8 = (C.c_void_p * B)()
        for b, n in enumerate(ns):
            m = max(n, 1)
            lab = np.full(m, -1, np.int8 if label8 else np.int32)
            order = np.zeros(m, np.int32) if want_order else None
            rs = np.zeros(URF_MAX_CHANNELS + 1, np.int32)
            if label8:
                l8[b] = lab.ctypes.data
            else:
                res[b].label = lab.ctypes.data_as(C.POINTER(C.c_int32))
            if want_order:
                res[b].order = order.ctypes.data_as(C.POINTER(C.c_int32))
            res[b].ring_start = rs.ctypes.data_as(C.POINTER(C.c_int32))
            keep.append((lab, order, rs))
        if point_step == 12 and (off_x, off_y, off_z) == (0, 4, 8) and off_intensity < 0:
            self._check(self.lib.urf_process_batch_xyz(self._ctx, ptrs, cn, B, res, l8 if label8 else None), "urf_process_batch_xyz")
        else:
            self._check(self.lib.urf_process_cloud2_batch(self._ctx, ptrs, cn, B, point_step, off_x, off_y, off_z, off_intensity, res,
                                                          l8 if label8 else None), "urf_process_cloud2_batch")
        out = []
        for b, n in enumerate(ns):
            lab, order, rs = keep[b]
            r = ScanResult()
            for f in ("status", "n_in", "n_roi", "n_rings", "n_order", "n_road", "n_curb", "n_vert", "flags"):
                setattr(r, f, int(getattr(res[b], f)))
            r.label = lab[:n].astype(np.int32)
            r.ring = None
            r.order = order[: r.n_order].copy() if want_order else None
            r.ring_start = rs[: r.n_rings + 1].copy()
            r.vert = np.ctypeslib.as_array(res[b].vert).reshape(URF_MAX_VERTS, 4)[: r.n_vert].copy()
            out.append(r)
        return out

    def filtered_cloud2(self, data: bytes | np.ndarray, n_points: int, point_step: int, off_x: int, off_y: int, off_z: int) -> ScanResult:
        """One scan from the raw `data` bytes of a sensor_msgs/PointCloud2 (unpacked on the device)."""
        raw = np.frombuffer(data, np.uint8) if isinstance(data, (bytes, bytearray)) else np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        m = max(n_points, 1)
        lab, ring, order = np.full(m, -1, np.int32), np.full(m, -1, np.int32), np.zeros(m, np.int32)
        rs = np.zeros(URF_MAX_CHANNELS + 1, np.int32)
        res = UrfResult()
        res.label = lab.ctypes.data_as(C.POINTER(C.c_int32)); res.ring = ring.ctypes.data_as(C.POINTER(C.c_int32))
        res.order = order.ctypes.data_as(C.POINTER(C.c_int32)); res.ring_start = rs.ctypes.data_as(C.POINTER(C.c_int32))
        self._check(self.lib.urf_process_cloud2(self._ctx, raw.ctypes.data, n_points, point_step, off_x, off_y, off_z, C.byref(res)),
                    "urf_process_cloud2")
        r = ScanResult()
        for f in ("status", "n_in", "n_roi", "n_rings", "n_order", "n_road", "n_curb", "n_vert", "flags"):
            setattr(r, f, int(getattr(res, f)))
        r.label, r.ring, r.order = lab[:n_points], ring[:n_points], order[: r.n_order].copy()
        r.ring_start = rs[: r.n_rings + 1].copy()
        r.vert = np.ctypeslib.as_array(res.vert).reshape(URF_MAX_VERTS, 4)[: r.n_vert].copy()
        return r

    def filtered_cloud2_packed(self, data, n_points: int, point_step: int, off_x: int, off_y: int, off_z: int,
                               off_intensity: int = -1, want_labels: bool = False):
        """One scan from raw PointCloud2 bytes; returns (ScanResult, clouds) where clouds maps "road" / "curb" / "roi" /
        "road_probably" to float32 arrays [count, 8] of 32-byte pcl::PointXYZI records packed on the device in the
        reference's emission order (include/urf.h urf_clouds). Labels / order are only fetched with want_labels."""
        raw = np.frombuffer(data, np.uint8) if isinstance(data, (bytes, bytearray)) else np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        m = max(n_points, 1)
        bufs = {k: np.empty((m, 8), np.float32) for k in ("road", "curb", "roi", "road_probably")}
        cl = UrfClouds()
        for k, a in bufs.items():
            setattr(cl, k, a.ctypes.data)
        rs = np.zeros(URF_MAX_CHANNELS + 1, np.int32)
        res = UrfResult()
        res.ring_start = rs.ctypes.data_as(C.POINTER(C.c_int32))
        lab = order = None
        if want_labels:
            lab, order = np.full(m, -1, np.int32), np.zeros(m, np.int32)
            res.label = lab.ctypes.data_as(C.POINTER(C.c_int32)); res.order = order.ctypes.data_as(C.POINTER(C.c_int32))
        self._check(self.lib.urf_process_cloud2_packed(self._ctx, raw.ctypes.data, n_points, point_step, off_x, off_y, off_z,
                                                       off_intensity, C.byref(res), C.byref(cl)), "urf_process_cloud2_packed")
        r = ScanResult()
        for f in ("status", "n_in", "n_roi", "n_rings", "n_order", "n_road", "n_curb", "n_vert", "flags"):
            setattr(r, f, int(getattr(res, f)))
        r.label = lab[:n_points] if want_labels else None
        r.order = order[: r.n_order].copy() if want_labels else None
        r.ring = None
        r.ring_start = rs[: r.n_rings + 1].copy()
        r.vert = np.ctypeslib.as_array(res.vert).reshape(URF_MAX_VERTS, 4)[: r.n_vert].copy()
        counts = dict(road=cl.n_road, curb=cl.n_curb, roi=cl.n_roi, road_probably=cl.n_road_probably)
        return r, {k: bufs[k][: counts[k]] for k in bufs}

    def filtered(self, cloud, **kw) -> ScanResult:
        """One Detector::filtered() call."""
        return self.filtered_batch([cloud], **kw)[0]

    def markers(self, result: ScanResult):
        """road_marker MarkerArray of a scan (keeps the reference's `ghostcount` state between scans)."""
        if not result.published:
            return []
        strips, self.ghostcount = build_markers(self.params, result.vert, self.ghostcount)
        return strips

    # diagnostics -------------------------------------------------------------------------------------------------
    def last_device_ms(self) -> float:
        return float(self.lib.urf_last_device_ms(self._ctx))

    def last_launch_count(self) -> int:
        return int(self.lib.urf_last_launch_count(self._ctx))

    def kernel_times(self, slot: int = 0) -> list[tuple[str, float]]:
        """(kernel name, device ms) of the call recorded in event slot `slot`; needs set_option(1, nslots) beforehand."""
        out = []
        for i in range(self.lib.urf_profile_count(self._ctx)):
            name, ms = C.c_char_p(), C.c_float()
            self._check(self.lib.urf_profile_get(self._ctx, slot, i, C.byref(name), C.byref(ms)), "urf_profile_get")
            out.append((name.value.decode(), float(ms.value)))
        return out

    def debug_fetch(self, scan: int, what: int, dtype, count: int) -> np.ndarray:
        a = np.zeros(max(count, 1), dtype)
        self._check(self.lib.urf_debug_fetch(self._ctx, scan, what, a.ctypes.data, a.nbytes if count else 0), "urf_debug_fetch")
        return a[:count]


class MultiGpuQueue:
    """One ingest stream over several GPUs (include/urf.h urf_mq, BASELINE config 4): a context + streaming queue per device,
    every scan goes to the device with the fewest scans in flight, results come back in submission order. Any number of
    producer threads, one consumer. `by_reference` submits hand the array to the library without a copy: keep it alive and
    unchanged until its result has come back."""

    def __init__(self, devices, max_points: int, slots_per_device: int = 8, max_batch: int = 4, params: UrfParams | None = None,
                 process_fn=None):
        self.lib = load_library()
        self._m = C.c_void_p()
        self.max_points = max_points
        self._cb = None
        if process_fn is not None:                      # tests: stand-in devices, no GPU
            self._cb = QUEUE_PROCESS_FN(process_fn)
            rc = self.lib.urf_mq_create_with(C.byref(self._m), self._cb, None, len(devices), max_points, slots_per_device, max_batch)
        else:
            dv = (C.c_int * len(devices))(*devices)
            rc = self.lib.urf_mq_create(C.byref(self._m), dv, len(devices), max_points, slots_per_device, max_batch,
                                        C.byref(params) if params is not None else None)
        if rc != URF_OK:
            raise UrfError(rc, "urf_mq_create", self.lib.urf_last_cuda_error(None).decode())
        self._keep = {}

    def set_params(self, prm: UrfParams):
        rc = self.lib.urf_mq_set_params(self._m, C.byref(prm))
        if rc != URF_OK:
            raise UrfError(rc, "urf_mq_set_params")

    def submit(self, cloud: np.ndarray, tag: int = 0, timeout_ms: int = -1, by_reference: bool = False) -> int:
        pts = np.ascontiguousarray(cloud, np.float32)
        if by_reference:
            self._keep[tag] = pts
            rc = self.lib.urf_mq_submit_ref(self._m, pts.ctypes.data, pts.shape[0], tag, timeout_ms)
        else:
            rc = self.lib.urf_mq_submit(self._m, pts.ctypes.data, pts.shape[0], tag, timeout_ms)
        if rc not in (URF_OK, URF_ERR_TIMEOUT, URF_ERR_CLOSED):
            raise UrfError(rc, "urf_mq_submit")
        return rc

    def next(self, timeout_ms: int = -1):
        """(tag, ScanResult) of the oldest scan, or None on timeout / when the closed queue is drained."""
        lab = np.full(self.max_points, -1, np.int32)
        res = UrfResult()
        res.label = lab.ctypes.data_as(C.POINTER(C.c_int32))
        tag = C.c_uint64()
        rc = self.lib.urf_mq_next(self._m, C.byref(tag), C.byref(res), timeout_ms)
        if rc in (URF_ERR_TIMEOUT, URF_ERR_CLOSED):
            return None
        if rc != URF_OK:
            raise UrfError(rc, "urf_mq_next")
        self._keep.pop(tag.value, None)
        r = ScanResult()
        for f in ("status", "n_in", "n_roi", "n_rings", "n_order", "n_road", "n_curb", "n_vert", "flags"):
            setattr(r, f, int(getattr(res, f)))
        r.label = lab[: r.n_in].copy()
        r.ring = r.order = r.ring_start = None
        r.vert = np.ctypeslib.as_array(res.vert).reshape(URF_MAX_VERTS, 4)[: r.n_vert].copy()
        return tag.value, r

    def stats(self) -> dict:
        st = UrfMqStats()
        self.lib.urf_mq_get_stats(self._m, C.byref(st))
        n = st.n_devices
        return dict(n_devices=n, pending=st.pending, submitted=list(st.submitted[:n]), delivered=list(st.delivered[:n]),
                    batches=list(st.batches[:n]), largest_batch=list(st.largest_batch[:n]))

    def close(self):
        if self._m:
            self.lib.urf_mq_close(self._m)

    def destroy(self):
        if self._m:
            self.lib.urf_mq_destroy(self._m)
            self._m = C.c_void_p()


class ScanQueue:
    """Streaming ingest (include/urf.h urf_queue): producers `submit` scans from any thread, one worker
    thread batches whatever is pending through the detector, `next` returns results in submission order. With
    `process_fn` (a Python callable with urf_process_batch's arguments) the queue runs without a GPU — tests only."""

    def __init__(self, detector: "Detector | None", max_points: int, slots: int = 8, max_batch: int = 4,
                 policy: int = URF_QUEUE_BLOCK, process_fn=None):
        self.lib = load_library()
        self._q = C.c_void_p()
        self.max_points = max_points
        self._cb = None
        if process_fn is not None:
            self._cb = QUEUE_PROCESS_FN(process_fn)
            rc = self.lib.urf_queue_create_with(C.byref(self._q), self._cb, None, max_points, slots, max_batch, policy)
        else:
            assert detector is not None and detector.max_batch >= max_batch and detector.max_points >= max_points
            self._det = detector          # keeps the ctx alive; nobody else may use it while the queue exists
            rc = self.lib.urf_queue_create(C.byref(self._q), detector._ctx, max_points, slots, max_batch, policy)
        if rc != URF_OK:
            raise UrfError(rc, "urf_queue_create")

    def submit(self, cloud: np.ndarray, tag: int = 0, timeout_ms: int = -1) -> int:
        """Returns URF_OK, URF_ERR_TIMEOUT or URF_ERR_CLOSED; raises on anything else."""
        pts = np.ascontiguousarray(cloud, np.float32)
        rc = self.lib.urf_queue_submit(self._q, pts.ctypes.data, pts.shape[0], tag, timeout_ms)
        if rc not in (URF_OK, URF_ERR_TIMEOUT, URF_ERR_CLOSED):
            raise UrfError(rc, "urf_queue_submit")
        return rc

    def next(self, timeout_ms: int = -1):
        """(tag, ScanResult) of the oldest finished scan, or None on timeout / when the closed queue is drained."""
        lab = np.full(self.max_points, -1, np.int32)
        res = UrfResult()
        res.label = lab.ctypes.data_as(C.POINTER(C.c_int32))
        tag = C.c_uint64()
        rc = self.lib.urf_queue_next(self._q, C.byref(tag), C.byref(res), timeout_ms)
        if rc in (URF_ERR_TIMEOUT, URF_ERR_CLOSED):
            return None
        if rc != URF_OK:
            raise UrfError(rc, "urf_queue_next")
        r = ScanResult()
        for f in ("status", "n_in", "n_roi", "n_rings", "n_order", "n_road", "n_curb", "n_vert", "flags"):
            setattr(r, f, int(getattr(res, f)))
        r.label = lab[: r.n_in].copy()
        r.ring = r.order = r.ring_start = None
        r.vert = np.ctypeslib.as_array(res.vert).reshape(URF_MAX_VERTS, 4)[: r.n_vert].copy()
        return int(tag.value), r

    def stats(self) -> dict:
        st = UrfQueueStats()
        self.lib.urf_queue_get_stats(self._q, C.byref(st))
        return {k: int(getattr(st, k)) for k, _ in UrfQueueStats._fields_ if k != "reserved"}

    def close(self):
        if self._q:
            self.lib.urf_queue_close(self._q)

    def destroy(self):
        if self._q:
            self.lib.urf_queue_destroy(self._q)
            self._q = C.c_void_p()
