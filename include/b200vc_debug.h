/* b200vc debug hooks -- NOT part of the production C-ABI (include/b200vc.h) and not exported by a default build.
 * Build with NVCC_FLAGS=-DB200VC_ENABLE_GDN_TRACE (video-compression_b200/build.py) to get:
 *
 *   CTA 0 of the following tcgen05 GDN launches stamps clock64() per tile and pipeline event into `device_buffer`
 *   (256 tiles x 16 slots of int64; NULL switches it off).  The buffer must outlive every launch made while it is
 *   installed: the hook is a process-global pointer (which is exactly why production builds do not carry it).
 *   Used by tools/gdn_trace.py.
 */
#ifndef B200VC_DEBUG_H_
#define B200VC_DEBUG_H_
#ifdef __cplusplus
extern "C" {
#endif
void b200vc_debug_set_gdn_trace(long long* device_buffer);
#ifdef __cplusplus
}
#endif
#endif
