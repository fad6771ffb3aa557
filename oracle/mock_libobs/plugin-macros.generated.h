/* Stand-in for the header CMake generates from src/plugin-macros.h.in in the
 * reference (test infrastructure; ENABLE_PROFILE/SHOW_ROI left undefined like
 * the default build, CMakeLists.txt:15-16). */
#pragma once
#define PLUGIN_NAME "obs-color-monitor"
#define PLUGIN_VERSION "0.9.5"
#define ID_PREFIX "net.nagater.obs-color-monitor."
#define blog(level, msg, ...) ((void)0)
