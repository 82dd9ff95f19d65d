"""
ctypes binding of the C ABI declared in ``include/reheatfunq_b200.h``.

The product library is ``librhfq_b200.so`` (CUDA, sm_100a), built in-tree by
``__graft_entry__.build()``.  There is no CPU fallback: if the library is
missing or no CUDA device is present, importing / calling fails loudly.
"""
import ctypes as C
import os

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_u64p = C.POINTER(C.c_uint64)

_HERE = os.path.dirname(os.path.abspath(__file__))
# RHFQ_LIB: alternative build of the same CUDA library (A/B tests of kernel variants)
LIB_PATH = os.environ.get("RHFQ_LIB") or os.path.join(_HERE, "librhfq_b200.so")

SYMBOLS = [
    "version", "device_count", "status_message", "batch_create", "batch_destroy", "cache_release",
    "batch_size", "batch_qmax", "batch_status", "batch_pdf", "batch_cdf", "batch_tail",
    "batch_tail_quantiles", "batch_get_locals", "batch_get_k", "batch_get_bli", "batch_saq_leaves",
    "batch_counters", "resilience_run", "resilience_shard_count", "resilience_shard_index",
    "bli_known_answer", "gcp_ln_phi", "gcp_kl_batch_max",
    "restricted_samples", "bootstrap_data_selection", "samples_count", "samples_index_count",
    "samples_get", "samples_destroy", "global_phmax",
]


class Api:
    """Typed function table over a loaded shared library."""

    def __init__(self, cdll, prefix="rhfq_"):
        self.cdll = cdll
        self.prefix = prefix
        g = lambda name: getattr(cdll, prefix + name)
        self.version = g("version")
        self.version.restype = C.c_char_p
        self.device_count = g("device_count")
        self.device_count.restype = C.c_int
        self.status_message = g("status_message")
        self.status_message.restype = C.c_char_p
        self.status_message.argtypes = [C.c_int]
        self.batch_create = g("batch_create")
        self.batch_create.restype = C.c_int
        self.batch_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                      C.c_void_p, C.c_int,
                                      C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_char_p, C.c_size_t]
        self.batch_destroy = g("batch_destroy")
        self.batch_destroy.restype = None
        self.batch_destroy.argtypes = [C.c_void_p]
        self.cache_release = g("cache_release")
        self.cache_release.restype = None
        self.cache_release.argtypes = [C.c_int]
        self.batch_size = g("batch_size")
        self.batch_size.restype = C.c_int64
        self.batch_size.argtypes = [C.c_void_p]
        self.batch_qmax = g("batch_qmax")
        self.batch_qmax.argtypes = [C.c_void_p, _dp]
        self.batch_status = g("batch_status")
        self.batch_status.argtypes = [C.c_void_p, _i32p]
        for name in ("batch_pdf", "batch_cdf", "batch_tail", "batch_tail_quantiles"):
            f = g(name)
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int,
                          C.c_void_p, C.c_char_p, C.c_size_t]
            setattr(self, name, f)
        self.batch_get_locals = g("batch_get_locals")
        self.batch_get_locals.argtypes = [C.c_void_p, C.c_int64, _dp]
        self.batch_get_k = g("batch_get_k")
        self.batch_get_k.restype = C.c_int64
        self.batch_get_k.argtypes = [C.c_void_p, C.c_int64, _dp, C.c_int64]
        self.batch_get_bli = g("batch_get_bli")
        self.batch_get_bli.restype = C.c_int64
        self.batch_get_bli.argtypes = [C.c_void_p, C.c_int64, C.c_int, _dp, C.c_int64]
        self.batch_saq_leaves = g("batch_saq_leaves")
        self.batch_saq_leaves.restype = C.c_int64
        self.batch_saq_leaves.argtypes = [C.c_void_p]
        self.batch_counters = g("batch_counters")
        self.batch_counters.argtypes = [C.c_void_p, _u64p]
        self.resilience_run = g("resilience_run")
        self.resilience_run.restype = C.c_int
        self.resilience_run.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int,
                                        C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_void_p,
                                        C.c_double, C.c_double, C.c_uint64, C.c_int, C.c_int,
                                        C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_char_p, C.c_size_t]
        self.resilience_shard_count = g("resilience_shard_count")
        self.resilience_shard_count.restype = C.c_int64
        self.resilience_shard_count.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int]
        self.resilience_shard_index = g("resilience_shard_index")
        self.resilience_shard_index.restype = C.c_int
        self.resilience_shard_index.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]
        self.bli_known_answer = g("bli_known_answer")
        self.bli_known_answer.restype = C.c_int
        self.bli_known_answer.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                          C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_int64,
                                          C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_char_p,
                                          C.c_size_t]
        self.gcp_ln_phi = g("gcp_ln_phi")
        self.gcp_ln_phi.restype = C.c_int
        self.gcp_ln_phi.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_char_p, C.c_size_t]
        self.gcp_kl_batch_max = g("gcp_kl_batch_max")
        self.gcp_kl_batch_max.restype = C.c_int
        self.gcp_kl_batch_max.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                          C.c_double, C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_char_p, C.c_size_t]
        self.restricted_samples = g("restricted_samples")
        self.restricted_samples.restype = C.c_int
        self.restricted_samples.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int64,
                                            C.c_double, C.c_int64, C.c_int64, C.c_void_p,
                                            C.c_char_p, C.c_size_t]
        self.bootstrap_data_selection = g("bootstrap_data_selection")
        self.bootstrap_data_selection.restype = C.c_int
        self.bootstrap_data_selection.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int64,
                                                  C.c_double, C.c_int64, C.c_void_p,
                                                  C.c_char_p, C.c_size_t]
        self.samples_count = g("samples_count")
        self.samples_count.restype = C.c_int64
        self.samples_count.argtypes = [C.c_void_p]
        self.samples_index_count = g("samples_index_count")
        self.samples_index_count.restype = C.c_int64
        self.samples_index_count.argtypes = [C.c_void_p]
        self.samples_get = g("samples_get")
        self.samples_get.restype = C.c_int
        self.samples_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        self.samples_destroy = g("samples_destroy")
        self.samples_destroy.restype = None
        self.samples_destroy.argtypes = [C.c_void_p]
        self.global_phmax = g("global_phmax")
        self.global_phmax.restype = C.c_int
        self.global_phmax.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, _dp,
                                      C.c_char_p, C.c_size_t]
        if prefix == "rhfq_":
            self.measure_fp64_peak = cdll.rhfq_measure_fp64_peak
            self.measure_fp64_peak.argtypes = [C.c_int, _dp, C.c_char_p, C.c_size_t]


_API = None


def api():
    """The product API.  Raises if the CUDA library has not been built."""
    global _API
    if _API is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "reheatfunq_b200: %s not found. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        _API = Api(C.CDLL(LIB_PATH), "rhfq_")
    return _API
