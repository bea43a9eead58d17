"""kvcached_b200 — B200-native (sm_100a) elastic KV-cache hot path behind the reference's API.

  vmm_ops            the reference's `kvcached.vmm_ops` surface (virtual KV tensors, map/unmap, PageAllocator)
  KVCacheManager     alloc / free / available_size / resize / trim / reserve on device-resident state
  block_pool         ElasticBlockPool: prefix-cache pinning (vLLM), ref counts / LRU on the device
  sglang_alloc       SGLang token-level and paged allocators: ids and slot indices stay on the device
  kv_ops             fused paged KV scatter / gather
  tp                 rank-consistent page mapping: one NCCL broadcast of the kernel's page list
  trace / replay     synthetic traces and the three replay modes

Everything computes in libkvb200.so (hand-written CUDA + the driver VMM loop); importing the
manager without the library raises.
"""
__version__ = "0.1.0"

from .kv_cache_manager import KVCachedConfigError, KVCacheManager  # noqa: E402,F401
from .block_pool import ElasticBlockPool  # noqa: E402,F401
