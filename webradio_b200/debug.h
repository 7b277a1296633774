// debug.h -- logging macros with the names the WebRadio glue code expects (LOG_DEBUG / LOG_INFO /
// LOG_ERROR, printf-style, to stderr).  WR_QUIET_DEBUG compiles LOG_DEBUG away.
#ifndef WEBRADIO_B200_DEBUG_H
#define WEBRADIO_B200_DEBUG_H

#include <cstdio>

#define WR_LOG_(level, fmt, ...) \
	std::fprintf(stderr, "[" level "] %s:%d: " fmt, __FILE__, __LINE__, ##__VA_ARGS__)

#ifdef WR_QUIET_DEBUG
#define LOG_DEBUG(fmt, ...) do { } while (0)
#else
#define LOG_DEBUG(fmt, ...) WR_LOG_("debug", fmt, ##__VA_ARGS__)
#endif
#define LOG_INFO(fmt, ...)  WR_LOG_("info", fmt, ##__VA_ARGS__)
#define LOG_ERROR(fmt, ...) WR_LOG_("error", fmt, ##__VA_ARGS__)

#endif
