/*
 * TSDRCodes.h -- status codes shared by the tsdr_* API and the tsdrplugin_* source-plugin ABI.
 *
 * Same names and values as the reference (TempestSDR/src/include/TSDRCodes.h:16-27): hosts such as the JNI glue
 * map these integers to exceptions (JavaGUI/jni/TSDRLibraryNDK.c:47-88), and plugins return them.
 */
#ifndef _TSDRCodes
#define _TSDRCodes

enum {
	TSDR_OK                       = 0,
	TSDR_ERR_PLUGIN               = 1,    /* plugin missing / not loaded / lacks a symbol */
	TSDR_WRONG_VIDEOPARAMS        = 2,    /* height, refresh rate or derived width unusable */
	TSDR_ALREADY_RUNNING          = 3,
	TSDR_PLUGIN_PARAMETERS_WRONG  = 4,
	TSDR_SAMPLE_RATE_WRONG        = 5,
	TSDR_CANNOT_OPEN_DEVICE       = 6,    /* also used here when no CUDA device / the GPU library fails */
	TSDR_INCOMPATIBLE_PLUGIN      = 7,    /* dlopen failed */
	TSDR_INVALID_PARAMETER        = 8,
	TSDR_INVALID_PARAMETER_VALUE  = 9,
	TSDR_NOT_RUNNING              = 10,
	TSDR_NOT_IMPLEMENTED          = 404
};

#endif
