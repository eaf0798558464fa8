/* Build shim for the reference tree (test infrastructure, see oracle/Makefile).
 * minivideo.h:35 includes this header unconditionally, but the reference only
 * generates it on WIN32 (minivideo/CMakeLists.txt:200-208). */
#ifndef MINIVIDEO_EXPORT_SHIM_H
#define MINIVIDEO_EXPORT_SHIM_H
#define minivideo_EXPORT
#endif
