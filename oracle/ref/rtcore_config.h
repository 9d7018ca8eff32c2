/* Hand-written stand-in for Embree's cmake-generated include/embree3/rtcore_config.h
 * (template: /root/reference/ext/embree/kernels/rtcore_config.h.in).  Values are what the reference's
 * top-level build selects: Embree 3.6.0 (ext/embree/CMakeLists.txt:17-19), static library, no API
 * namespace, two instance levels (/root/reference/CMakeLists.txt:20,25).  Test infrastructure only. */
#pragma once
#define RTC_VERSION_MAJOR 3
#define RTC_VERSION_MINOR 6
#define RTC_VERSION_PATCH 0
#define RTC_VERSION 30600
#define RTC_VERSION_STRING "3.6.0"
#define RTC_MAX_INSTANCE_LEVEL_COUNT 2
#define EMBREE_STATIC_LIB
#define RTC_NAMESPACE_BEGIN
#define RTC_NAMESPACE_END
#define RTC_NAMESPACE_OPEN
#if defined(__cplusplus)
#define RTC_API_EXTERN_C extern "C"
#else
#define RTC_API_EXTERN_C
#endif
#define RTC_API_IMPORT RTC_API_EXTERN_C
#define RTC_API_EXPORT RTC_API_EXTERN_C
#if defined(RTC_EXPORT_API)
#define RTC_API RTC_API_EXPORT
#else
#define RTC_API RTC_API_IMPORT
#endif
