/* A minimal C host, as INTEGRATION.md section A describes it: compiled against include/TSDRLibrary.h, linked with
 * libTSDRLibrary.a + libtsdrgpu.so.  It drives the public API the way the JNI glue does (TSDRLibraryNDK.c:168-177,293-337):
 * init -> setresolution -> loadplugin -> readasync -> free, and prints what happened.  Exit code 0 = the calls behaved as
 * the contract says for the machine it ran on (with a GPU: frames arrive; without: TSDR_CANNOT_OPEN_DEVICE, loudly). */
#include "TSDRLibrary.h"
#include "TSDRCodes.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

static volatile int frames = 0, finished = 0;
static tsdr_lib_t *lib;

static void on_frame(float *buf, int width, int height, void *ctx) { (void) buf; (void) ctx; if (width > 0 && height > 0) frames++; }
static void on_value(int id, double a, double b, void *ctx) { (void) id; (void) a; (void) b; (void) ctx; }
static void on_plot(int id, int offset, double *values, int size, uint32_t samplerate, void *ctx) {
	(void) id; (void) offset; (void) values; (void) size; (void) samplerate; (void) ctx;
}
static void *stopper(void *arg) {
	(void) arg;
	for (int i = 0; i < 400 && frames < 3 && !finished; i++) usleep(50 * 1000);
	if (!finished) tsdr_stop(lib);
	return NULL;
}

int main(int argc, char **argv) {
	if (argc < 3) { fprintf(stderr, "usage: host_smoke <plugin.so> <plugin params>\n"); return 2; }
	tsdr_init(&lib, on_value, on_plot, NULL);
	if (!lib) return 3;
	if (tsdr_setresolution(lib, 125, 60.0) != TSDR_OK) return 4;
	if (tsdr_isrunning(lib) != 0) return 5;
	if (tsdr_stop(lib) != TSDR_OK) return 6;                       /* stopping an idle library is fine (TSDRLibrary.c:213) */
	int rc = tsdr_loadplugin(lib, argv[1], argv[2]);
	if (rc != TSDR_OK) { printf("loadplugin rc=%d text=%s\n", rc, tsdr_getlasterrortext(lib)); tsdr_free(&lib); return lib ? 7 : 10; }
	pthread_t th;
	pthread_create(&th, NULL, stopper, NULL);
	rc = tsdr_readasync(lib, on_frame, NULL);                       /* blocks until tsdr_stop or an error */
	finished = 1;
	const char *text = tsdr_getlasterrortext(lib);
	printf("readasync rc=%d frames=%d text=%s\n", rc, frames, text ? text : "(null)");
	pthread_join(th, NULL);
	tsdr_free(&lib);
	if (lib != NULL) return 8;                                      /* tsdr_free NULLs the handle (TSDRLibrary.c:96-108) */
	if (rc == TSDR_OK) return frames >= 3 ? 0 : 9;
	return rc == TSDR_CANNOT_OPEN_DEVICE ? 0 : 11;
}
