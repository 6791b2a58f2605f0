"""ctypes binding of the C ABI in ``include/gqe.h`` (libgqe_b200.so).

The shared library is built in-tree by ``build()`` (``nvcc`` for sm_100a) and
loaded from the package directory.  There is no fallback: if the library is
missing, or no sm_100 device is present, every entry point raises.
"""
import ctypes as C
import os
import subprocess
import sys

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_NAME = "libgqe_b200.so"
LIB_PATH = os.environ.get("GQE_LIB_PATH") or os.path.join(_PKG, LIB_NAME)   # (override: kernel A/B experiments)
CSRC = os.path.join(_PKG, "csrc")
BUILD_DIR = os.path.join(_PKG, "build")


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    return sorted(hs) + [os.path.join(_ROOT, "include", "gqe.h")]


GQE_MAX_ANCHORS = 3
GQE_MAX_RELS = 3

STRUCTURE_ID = {"1-chain": 0, "2-chain": 1, "3-chain": 2, "2-inter": 3, "3-inter": 4,
                "3-inter_chain": 5, "3-chain_inter": 6}
DECODER_ID = {"bilinear": 0, "transe": 1, "bilinear-diag": 2}
INTER_ID = {"mean": 0, "min": 1, "mean-simple": 2, "min-simple": 3}
PRECISION_ID = {"bf16x3": 0, "fp32": 1}
COMPOSE_ID = {"off": 0, "auto": 1, "always": 2}
ABI_VERSION = 6


GQE_ERR_INDEX = -6


class GqeError(RuntimeError):
    def __init__(self, code, message):
        RuntimeError.__init__(self, "gqe error %d: %s" % (code, message))
        self.code = code


class GqeIndexError(GqeError, KeyError, IndexError):
    """A node id that is not in the bound node map (the reference raises KeyError from its
    node_maps dict, bio/data_utils.py:21) or a row outside its table (nn.Embedding's
    IndexError).  Catchable as either."""

    def __str__(self):
        return RuntimeError.__str__(self)


class Plan(C.Structure):
    _fields_ = [("structure", C.c_int32), ("target_mode", C.c_int32),
                ("anchor_mode", C.c_int32 * GQE_MAX_ANCHORS), ("inter_mode", C.c_int32),
                ("rel", C.c_int32 * GQE_MAX_RELS)]

    def as_tuple(self):
        return (self.structure, self.target_mode, tuple(self.anchor_mode), self.inter_mode, tuple(self.rel))


class AdamHyper(C.Structure):
    """gqe_adam: torch.optim.Adam's hyper-parameters (defaults as torch's)."""
    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float)]

    def __init__(self, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
        C.Structure.__init__(self, float(lr), float(beta1), float(beta2), float(eps))


class Segment(C.Structure):
    _fields_ = [("plan", Plan), ("query_begin", C.c_int64), ("query_end", C.c_int64)]


class StoreSliceC(C.Structure):
    """gqe_store_slice: one formula's slice of a device-resident query store (device pointers)."""
    _fields_ = [("anchors", C.c_void_p), ("targets", C.c_void_p), ("neg_ptr", C.c_void_p), ("negs", C.c_void_p),
                ("block_queries", C.c_int64), ("start", C.c_int64), ("pool_size", C.c_int64)]


# name -> (restype, argtypes); every symbol include/gqe.h declares.
_P = C.c_void_p
_SIGNATURES = {
    "gqe_abi_version": (C.c_int, []),
    "gqe_create": (C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    "gqe_destroy": (None, [_P]),
    "gqe_set_stream": (C.c_int, [_P, _P]),
    "gqe_set_precision": (C.c_int, [_P, C.c_int32]),
    "gqe_get_precision": (C.c_int, [_P]),
    "gqe_set_compose": (C.c_int, [_P, C.c_int32]),
    "gqe_set_weight_cache": (C.c_int, [_P, C.c_int32]),
    "gqe_invalidate_weights": (C.c_int, [_P]),
    "gqe_weight_prep_count": (C.c_int64, [_P]),
    "gqe_bind_node_maps": (C.c_int, [_P, C.c_int32, C.POINTER(_P), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "gqe_index_error": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "gqe_last_error": (C.c_char_p, [_P]),
    "gqe_launch_count": (C.c_int64, [_P]),
    "gqe_debug_set_phase_log": (C.c_int, [_P, _P, C.c_int64]),
    "gqe_debug_score_col_src": (C.c_int, [C.c_int]),
    "gqe_bind_tables": (C.c_int, [_P, C.c_int32, C.POINTER(_P), C.POINTER(C.c_int64), C.c_int32]),
    "gqe_bind_relations": (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(_P), C.c_int32]),
    "gqe_bind_intersection": (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(_P), C.POINTER(_P), C.c_int32, C.c_int32]),
    "gqe_score_device": (C.c_int, [_P, C.POINTER(Plan), C.c_int64, _P, C.c_int64, _P, _P, _P]),
    "gqe_margin_loss_device": (C.c_int, [_P, C.POINTER(Plan), C.c_int64, _P, _P, C.c_float, _P, _P]),
    "gqe_score_grouped_device": (C.c_int, [_P, C.POINTER(Segment), C.c_int32, C.c_int64, _P, _P, C.c_int32, _P,
                                           C.c_float, _P]),
    "gqe_score_nodes_device": (C.c_int, [_P, C.POINTER(Plan), C.c_int64, _P, C.c_int64, _P, _P, _P]),
    "gqe_margin_loss_nodes_device": (C.c_int, [_P, C.POINTER(Plan), C.c_int64, _P, _P, C.c_float, _P, _P]),
    "gqe_score_grouped_nodes_device": (C.c_int, [_P, C.POINTER(Segment), C.c_int32, C.c_int64, _P, _P, C.c_int32, _P,
                                                 C.c_float, _P]),
    "gqe_margin_loss_store_device": (C.c_int, [_P, C.POINTER(Segment), C.c_int32, C.POINTER(StoreSliceC), C.c_uint64,
                                               C.c_float, _P, _P, _P]),
    "gqe_score_nodes_host": (C.c_int, [_P, C.POINTER(Plan), C.c_int64, _P, C.c_int64, _P, _P, _P]),
    "gqe_margin_loss_nodes_host": (C.c_int, [_P, C.POINTER(Plan), C.c_int64, _P, _P, C.c_float, _P, _P]),
    "gqe_score_grouped_nodes_host": (C.c_int, [_P, C.POINTER(Segment), C.c_int32, C.c_int64, _P, _P, C.c_int32, _P,
                                               C.c_float, _P]),
    "gqe_score_host": (C.c_int, [_P, C.POINTER(Plan), C.c_int64, _P, C.c_int64, _P, _P, _P]),
    "gqe_margin_loss_host": (C.c_int, [_P, C.POINTER(Plan), C.c_int64, _P, _P, C.c_float, _P, _P]),
    "gqe_score_grouped_host": (C.c_int, [_P, C.POINTER(Segment), C.c_int32, C.c_int64, _P, _P, C.c_int32, _P,
                                         C.c_float, _P]),
    "gqe_encode_device": (C.c_int, [_P, C.c_int32, C.c_int64, _P, _P]),
    "gqe_project_device": (C.c_int, [_P, C.c_int32, C.c_int64, _P, _P]),
    "gqe_path_score_device": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_int32), C.c_int64, _P, _P, C.c_int32, _P]),
    "gqe_intersect_device": (C.c_int, [_P, C.c_int32, C.c_int64, _P, _P, _P, _P]),
    "gqe_cosine_device": (C.c_int, [_P, C.c_int32, C.c_int64, _P, _P, _P]),
    "gqe_matmul_device": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int64, _P, _P]),
    "gqe_matmul_wgrad_device": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int64, _P, _P, _P]),
    "gqe_rowsum_device": (C.c_int, [_P, C.c_int32, C.c_int64, _P, _P, _P]),
    "gqe_aggregate_device": (C.c_int, [_P, C.c_int32, C.c_int64, _P, _P, _P, C.c_int32, C.c_int32, _P]),
    "gqe_aggregate_bwd_device": (C.c_int, [_P, C.c_int32, C.c_int64, _P, _P, _P, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "gqe_dot_device": (C.c_int, [_P, C.c_int32, C.c_int64, _P, _P, _P]),
    "gqe_cosine_bwd_device": (C.c_int, [_P, C.c_int32, C.c_int64, _P, _P, _P, C.c_int32, _P, _P]),
    "gqe_encode_bwd_device": (C.c_int, [_P, C.c_int32, C.c_int64, _P, _P, _P]),
    "gqe_encode_bwd_rows_device": (C.c_int, [_P, C.c_int32, C.c_int64, _P, _P, _P]),
    "gqe_adam_rows_device": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int64, _P, _P, C.c_int32, C.c_float,
                                       C.c_float, C.c_float, C.c_float]),
    "gqe_train_step_device": (C.c_int, [_P, C.POINTER(Plan), C.c_int64, _P, _P, C.c_float, _P, _P]),
    "gqe_train_step_nodes_device": (C.c_int, [_P, C.POINTER(Plan), C.c_int64, _P, _P, C.c_float, _P, _P]),
    "gqe_train_step_host": (C.c_int, [_P, C.POINTER(Plan), C.c_int64, _P, _P, C.c_float, _P, _P]),
    "gqe_train_step_nodes_host": (C.c_int, [_P, C.POINTER(Plan), C.c_int64, _P, _P, C.c_float, _P, _P]),
    "gqe_train_backward_device": (C.c_int, [_P, C.POINTER(Plan), C.c_int64, _P, _P, C.c_float, C.c_float, _P, _P]),
    "gqe_train_backward_nodes_device": (C.c_int, [_P, C.POINTER(Plan), C.c_int64, _P, _P, C.c_float, C.c_float, _P, _P]),
    "gqe_train_backward_host": (C.c_int, [_P, C.POINTER(Plan), C.c_int64, _P, _P, C.c_float, C.c_float, _P, _P]),
    "gqe_train_backward_nodes_host": (C.c_int, [_P, C.POINTER(Plan), C.c_int64, _P, _P, C.c_float, C.c_float, _P, _P]),
    "gqe_train_apply": (C.c_int, [_P, _P]),
    "gqe_train_flush": (C.c_int, [_P]),
    "gqe_train_reset": (C.c_int, [_P]),
    "gqe_train_steps": (C.c_int64, [_P, C.c_int32]),
    "gqe_segment_mean_device": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int64, _P, _P, _P]),
    "gqe_linear_device": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int64, _P, C.c_int32, _P]),
    "gqe_ipc_export": (C.c_int, [_P, _P, C.c_char_p, C.POINTER(C.c_int64)]),
    "gqe_ipc_open": (C.c_int, [_P, C.c_char_p, C.c_int64, C.POINTER(_P)]),
    "gqe_ipc_close": (C.c_int, [_P, _P]),
    "gqe_gather_rows_device": (C.c_int, [_P, C.c_int32, C.c_int64, _P, _P]),
}
IPC_HANDLE_BYTES = 64
EXPORTED_SYMBOLS = tuple(sorted(_SIGNATURES))

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > built for p in _sources() + _headers())


def build(force=False, verbose=False, jobs=None):
    """Compile every csrc/*.cu for sm_100a (one nvcc per translation unit, in
    parallel) and link them into the in-tree shared library."""
    if not force and not needs_build():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(BUILD_DIR, exist_ok=True)
    newest_header = max(os.path.getmtime(h) for h in _headers())

    def compile_one(src):
        obj = os.path.join(BUILD_DIR, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), newest_header):
            return obj, ""
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed (%d):\n%s\n%s" % (proc.returncode, " ".join(cmd), proc.stdout))
        return obj, proc.stdout

    srcs = _sources()
    with ThreadPoolExecutor(max_workers=jobs or min(len(srcs), os.cpu_count() or 4)) as pool:
        results = list(pool.map(compile_one, srcs))
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + [o for o, _ in results]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed (%d):\n%s\n%s" % (proc.returncode, " ".join(cmd), proc.stdout))
    return LIB_PATH


_lib = None


def load():
    """Load libgqe_b200.so and declare every prototype.  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "%s is not built; run `python -c \"import __graft_entry__ as g; g.build()\"` "
            "(there is no CPU fallback for the CUDA path)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.gqe_abi_version() != ABI_VERSION:
        raise RuntimeError("libgqe_b200.so ABI version mismatch")
    _lib = lib
    return lib


def _ptr_array(ptrs):
    arr = (_P * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr


class Context(object):
    """A gqe_ctx: one device, one stream, one set of bound parameters."""

    def __init__(self, device=0, stream=None):
        self._lib = load()
        handle = _P()
        rc = self._lib.gqe_create(int(device), _P(stream or 0), C.byref(handle))
        if rc != 0:
            raise GqeError(rc, self._lib.gqe_last_error(None).decode())
        self._h = handle
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gqe_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = self._lib.gqe_last_error(self._h).decode()
            raise (GqeIndexError if rc == GQE_ERR_INDEX else GqeError)(rc, msg)

    def set_stream(self, stream):
        self._check(self._lib.gqe_set_stream(self._h, _P(stream or 0)))

    def set_precision(self, precision):
        """"bf16x3" (tensor cores, default) or "fp32" (exact CUDA-core FMA)."""
        self._check(self._lib.gqe_set_precision(self._h, PRECISION_ID[precision]))

    def get_precision(self):
        code = int(self._lib.gqe_get_precision(self._h))
        return [k for k, v in PRECISION_ID.items() if v == code][0]

    def set_compose(self, mode):
        """Operator pre-composition on the tensor-core path: "off", "auto" (default), "always"."""
        self._check(self._lib.gqe_set_compose(self._h, COMPOSE_ID[mode]))

    def launch_count(self):
        return int(self._lib.gqe_launch_count(self._h))

    def set_weight_cache(self, on):
        """Cache packed / pre-multiplied operator matrices across calls (default on); with the
        cache on, call ``invalidate_weights`` after changing a bound matrix in place."""
        self._check(self._lib.gqe_set_weight_cache(self._h, 1 if on else 0))

    def invalidate_weights(self):
        self._check(self._lib.gqe_invalidate_weights(self._h))

    def weight_prep_count(self):
        return int(self._lib.gqe_weight_prep_count(self._h))

    def bind_node_maps(self, lut_ptrs, bases, lens):
        """Per mode: device pointer of the int32 lut (0: affine map), first node id, lut length."""
        n = len(bases)
        lut = _ptr_array(lut_ptrs) if lut_ptrs is not None else None
        self._check(self._lib.gqe_bind_node_maps(self._h, n, lut, (C.c_int64 * n)(*[int(b) for b in bases]),
                                                 (C.c_int64 * n)(*[int(x) for x in lens])))

    def index_error(self):
        """Synchronise and raise GqeIndexError if a kernel of this context saw a bad index."""
        k, m, v = C.c_int32(), C.c_int32(), C.c_int64()
        self._check(self._lib.gqe_index_error(self._h, C.byref(k), C.byref(m), C.byref(v)))

    def debug_set_phase_log(self, log_ptr, n_tiles):
        self._check(self._lib.gqe_debug_set_phase_log(self._h, _P(log_ptr or 0), int(n_tiles)))

    # -- binding ------------------------------------------------------------
    def bind_tables(self, table_ptrs, rows, d):
        n = len(table_ptrs)
        rows_arr = (C.c_int64 * n)(*[int(r) for r in rows])
        self._check(self._lib.gqe_bind_tables(self._h, n, _ptr_array(table_ptrs), rows_arr, int(d)))

    def bind_relations(self, decoder, param_ptrs, d):
        self._check(self._lib.gqe_bind_relations(self._h, int(decoder), len(param_ptrs), _ptr_array(param_ptrs), int(d)))

    def bind_intersection(self, inter, pre_ptrs, post_ptrs, d, d_expand=None):
        n = len(pre_ptrs or [])
        pre = _ptr_array(pre_ptrs) if n else None
        post = _ptr_array(post_ptrs) if n else None
        self._check(self._lib.gqe_bind_intersection(self._h, int(inter), n, pre, post, int(d),
                                                    int(d if d_expand is None else d_expand)))

    # -- fused path, device buffers (raw device pointers as ints) -----------------
    # ``nodes=True`` selects the "_nodes" twin: the index arrays hold node ids, mapped to rows
    # inside the kernels through the bound node maps.
    def score_device(self, plan, n_queries, anchor_rows, n_pairs, target_rows, target_offsets, out_scores, nodes=False):
        fn = self._lib.gqe_score_nodes_device if nodes else self._lib.gqe_score_device
        self._check(fn(self._h, C.byref(plan), n_queries, anchor_rows, n_pairs, target_rows, target_offsets, out_scores))

    def margin_loss_device(self, plan, n_queries, anchor_rows, pair_rows, margin, out_loss, out_scores=None, nodes=False):
        fn = self._lib.gqe_margin_loss_nodes_device if nodes else self._lib.gqe_margin_loss_device
        self._check(fn(self._h, C.byref(plan), n_queries, anchor_rows, pair_rows, float(margin), out_loss, out_scores))

    def score_grouped_device(self, segments, n_queries_total, anchor_rows, target_rows, targets_per_query,
                             out_scores, margin=1.0, out_loss=None, nodes=False):
        fn = self._lib.gqe_score_grouped_nodes_device if nodes else self._lib.gqe_score_grouped_device
        self._check(fn(self._h, segments, len(segments), n_queries_total, anchor_rows, target_rows, targets_per_query,
                       out_scores, float(margin), out_loss))

    def margin_loss_store_device(self, segments, slices, seed, margin, out_loss, out_scores=None, out_pairs=None):
        """segments: gqe_segment array; slices: StoreSliceC array (same length); device pointers out."""
        self._check(self._lib.gqe_margin_loss_store_device(self._h, segments, len(segments), slices, int(seed) & (2 ** 64 - 1),
                                                           float(margin), out_loss, out_scores, out_pairs))

    # -- fused path, host buffers (numpy arrays) -----------------------------------
    def score_host(self, plan, anchor_rows, target_rows, target_offsets, out_scores, nodes=False):
        nq = anchor_rows.shape[1]
        off = None if target_offsets is None else target_offsets.ctypes.data
        fn = self._lib.gqe_score_nodes_host if nodes else self._lib.gqe_score_host
        self._check(fn(self._h, C.byref(plan), nq, anchor_rows.ctypes.data, target_rows.size, target_rows.ctypes.data,
                       off, out_scores.ctypes.data))

    def margin_loss_host(self, plan, anchor_rows, pair_rows, margin, out_loss, out_scores=None, nodes=False):
        nq = anchor_rows.shape[1]
        sc = None if out_scores is None else out_scores.ctypes.data
        fn = self._lib.gqe_margin_loss_nodes_host if nodes else self._lib.gqe_margin_loss_host
        self._check(fn(self._h, C.byref(plan), nq, anchor_rows.ctypes.data, pair_rows.ctypes.data, float(margin),
                       out_loss.ctypes.data, sc))

    def score_grouped_host(self, segments, anchor_rows, target_rows, targets_per_query, out_scores, margin=1.0,
                           out_loss=None, nodes=False):
        nq = anchor_rows.shape[1]
        sc = None if out_scores is None else out_scores.ctypes.data
        ls = None if out_loss is None else out_loss.ctypes.data
        fn = self._lib.gqe_score_grouped_nodes_host if nodes else self._lib.gqe_score_grouped_host
        self._check(fn(self._h, segments, len(segments), nq, anchor_rows.ctypes.data, target_rows.ctypes.data,
                       targets_per_query, sc, float(margin), ls))

    # -- operator level ---------------------------------------------------------------
    def encode_device(self, mode, n, rows, out):
        self._check(self._lib.gqe_encode_device(self._h, mode, n, rows, out))

    def project_device(self, rel, n, src, out):
        self._check(self._lib.gqe_project_device(self._h, rel, n, src, out))

    def path_score_device(self, rels, n, embeds1, embeds2, mutate, out):
        arr = (C.c_int32 * len(rels))(*rels)
        self._check(self._lib.gqe_path_score_device(self._h, len(rels), arr, n, embeds1, embeds2, int(mutate), out))

    def intersect_device(self, mode, n, e1, e2, e3, out):
        self._check(self._lib.gqe_intersect_device(self._h, mode, n, e1, e2, e3, out))

    def cosine_device(self, d, n, x, y, out):
        self._check(self._lib.gqe_cosine_device(self._h, d, n, x, y, out))

    # -- differentiable operator surface (raw device pointers as ints) ---------------------------
    def matmul_device(self, w, transpose, d, n, src, out):
        self._check(self._lib.gqe_matmul_device(self._h, w, int(transpose), d, n, src, out))

    def matmul_wgrad_device(self, transpose, d, n, gy, x, gw):
        self._check(self._lib.gqe_matmul_wgrad_device(self._h, int(transpose), d, n, gy, x, gw))

    def rowsum_device(self, d, n, gy, x, gv):
        self._check(self._lib.gqe_rowsum_device(self._h, d, n, gy, x, gv))

    def aggregate_device(self, d, n, e1, e2, e3, relu, use_min, out):
        self._check(self._lib.gqe_aggregate_device(self._h, d, n, e1, e2, e3, int(relu), int(use_min), out))

    def aggregate_bwd_device(self, d, n, e1, e2, e3, relu, use_min, gout, g1, g2, g3):
        self._check(self._lib.gqe_aggregate_bwd_device(self._h, d, n, e1, e2, e3, int(relu), int(use_min), gout, g1, g2, g3))

    def dot_device(self, d, n, x, y, out):
        self._check(self._lib.gqe_dot_device(self._h, d, n, x, y, out))

    def cosine_bwd_device(self, d, n, x, y, gout, raw_dot, gx, gy):
        self._check(self._lib.gqe_cosine_bwd_device(self._h, d, n, x, y, gout, int(raw_dot), gx, gy))

    def encode_bwd_device(self, mode, n, rows, gout, gtable):
        self._check(self._lib.gqe_encode_bwd_device(self._h, int(mode), n, rows, gout, gtable))

    def encode_bwd_rows_device(self, mode, n, rows, gout, grad_rows):
        self._check(self._lib.gqe_encode_bwd_rows_device(self._h, int(mode), n, rows, gout, grad_rows))

    def adam_rows_device(self, table, exp_avg, exp_avg_sq, last_step, table_rows, d, n, rows, grad_rows, step, lr,
                         beta1, beta2, eps):
        self._check(self._lib.gqe_adam_rows_device(self._h, table, exp_avg, exp_avg_sq, last_step, int(table_rows), int(d),
                                                   int(n), rows, grad_rows, int(step), float(lr), float(beta1),
                                                   float(beta2), float(eps)))

    # -- native training step ------------------------------------------------------------
    def train_step_device(self, plan, n_queries, anchors, pairs, margin, hyper, out_loss, nodes=False):
        """hyper: an AdamHyper; anchors / pairs / out_loss: device pointers."""
        fn = self._lib.gqe_train_step_nodes_device if nodes else self._lib.gqe_train_step_device
        self._check(fn(self._h, C.byref(plan), int(n_queries), anchors, pairs, float(margin), C.byref(hyper), out_loss))

    def train_step_host(self, plan, n_queries, anchors, pairs, margin, hyper, nodes=False):
        """anchors / pairs: host pointers (ideally pinned); -> the loss before the update."""
        fn = self._lib.gqe_train_step_nodes_host if nodes else self._lib.gqe_train_step_host
        out = C.c_float(0.0)
        self._check(fn(self._h, C.byref(plan), int(n_queries), anchors, pairs, float(margin), C.byref(hyper), C.byref(out)))
        return out.value

    def train_backward_host(self, plan, n_queries, anchors, pairs, margin, weight, hyper, nodes=False):
        """forward + backward, gradients x weight accumulated, no update; -> the batch's own loss."""
        fn = self._lib.gqe_train_backward_nodes_host if nodes else self._lib.gqe_train_backward_host
        out = C.c_float(0.0)
        self._check(fn(self._h, C.byref(plan), int(n_queries), anchors, pairs, float(margin), float(weight), C.byref(hyper),
                       C.byref(out)))
        return out.value

    def train_apply(self, hyper):
        self._check(self._lib.gqe_train_apply(self._h, C.byref(hyper)))

    def train_flush(self):
        self._check(self._lib.gqe_train_flush(self._h))

    def train_reset(self):
        self._check(self._lib.gqe_train_reset(self._h))

    def train_steps(self, mode):
        return int(self._lib.gqe_train_steps(self._h, int(mode)))

    def segment_mean_device(self, src, n_src, d, n, ptr, cols, out):
        self._check(self._lib.gqe_segment_mean_device(self._h, src, int(n_src), int(d), int(n), ptr, cols, out))

    def linear_device(self, w, m, k, n, x, relu, out):
        self._check(self._lib.gqe_linear_device(self._h, w, int(m), int(k), int(n), x, int(relu), out))

    # -- node-type-sharded tables ---------------------------------------------------------
    def ipc_export(self, dev_ptr):
        """-> (handle bytes[64], byte offset) of a device pointer of this process."""
        buf = C.create_string_buffer(IPC_HANDLE_BYTES)
        off = C.c_int64()
        self._check(self._lib.gqe_ipc_export(self._h, _P(dev_ptr), buf, C.byref(off)))
        return buf.raw, int(off.value)

    def ipc_open(self, handle, offset):
        """Map a peer rank's shard; returns the peer device pointer (int)."""
        out = _P()
        self._check(self._lib.gqe_ipc_open(self._h, bytes(handle), int(offset), C.byref(out)))
        return int(out.value)

    def ipc_close(self, peer_ptr):
        self._check(self._lib.gqe_ipc_close(self._h, _P(peer_ptr)))

    def gather_rows_device(self, mode, n, rows, out):
        self._check(self._lib.gqe_gather_rows_device(self._h, int(mode), int(n), rows, out))


def make_segments(items):
    """[(Plan, begin, end), ...] -> ctypes array of gqe_segment."""
    arr = (Segment * len(items))()
    for i, (plan, b, e) in enumerate(items):
        arr[i].plan = plan
        arr[i].query_begin = int(b)
        arr[i].query_end = int(e)
    return arr
