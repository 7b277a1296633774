// tests/harness/host_threads.cxx -- TEST DRIVER: the drop-in blocks under ThreadSanitizer (make
// tsan-check; stand-in back-end).  One thread runs the pipeline the way src/main.cxx:114-115 does,
// two others do what libmicrohttpd's connection threads do through the REST handlers (reference
// src/web/receiverhandler.cxx:113-140, waterfallhandler.cxx:56-61): setIF / setModeString /
// setPassband and their getters, getSpectrum -- with no locking on the caller's side, as in the
// reference.
#include <atomic>
#include <cstdio>
#include <thread>
#include <vector>
extern "C" {
void wrh_set_quiet(int);
void *wrh_graph_create(unsigned fs, unsigned block_frames);
int wrh_graph_add_receiver(void *h, int if_hz, unsigned ch_passband, unsigned ch_rate, unsigned ch_decim, int mode,
		unsigned au_passband, unsigned au_rate, unsigned au_decim, unsigned capture_mask);
int wrh_graph_add_spectrum(void *h, unsigned fft_size);
int wrh_graph_start(void *h);
int wrh_graph_run(void *h, const float *iq);
int wrh_graph_set_if(void *h, int rx, int hz);
int wrh_graph_set_mode(void *h, int rx, const char *mode);
int wrh_graph_get_mode(void *h, int rx);
int wrh_graph_set_passband(void *h, int rx, int which, unsigned hz);
int wrh_graph_get_taps(void *h, int rx, int which, float *out, unsigned cap);
int wrh_graph_spectrum(void *h, float *db);
void wrh_graph_destroy(void *h);
}
int main()
{
	wrh_set_quiet(0);
	const unsigned fs = 2400000, F = 20000;
	void *g = wrh_graph_create(fs, F);
	for (int i = 0; i < 4; i++)
		wrh_graph_add_receiver(g, 10000 * i - 7, 80000, 240000, 0, i % 4, 8000, 48000, 0, i == 3 ? 0xF : 0x8);
	wrh_graph_add_spectrum(g, 512);
	if (wrh_graph_start(g)) { fprintf(stderr, "start failed\n"); return 1; }
	std::atomic<bool> stop(false);
	std::thread http1([&] {
		const char *modes[] = { "AM", "FM", "USB", "LSB" };
		std::vector<float> taps(64);
		for (unsigned k = 0; !stop.load(); k++) {
			wrh_graph_set_if(g, k % 4, (int)(k * 7919 % 2000000) - 1000000);
			wrh_graph_set_mode(g, (k + 1) % 4, modes[k % 4]);
			wrh_graph_get_mode(g, k % 4);
			wrh_graph_set_passband(g, (k + 2) % 4, k & 1, 20000 + (k % 50) * 4000);
			wrh_graph_get_taps(g, k % 4, k & 1, taps.data(), 64);
		}
	});
	std::thread http2([&] {
		std::vector<float> db(512);
		while (!stop.load())
			wrh_graph_spectrum(g, db.data());
	});
	std::vector<float> iq(2 * F);
	unsigned s = 99;
	int rc = 0;
	for (int b = 0; b < 60 && !rc; b++) {
		for (size_t k = 0; k < iq.size(); k++) { s = s * 1664525u + 1013904223u; iq[k] = ((float)(s >> 24) - 128.0f) / 128.0f; }
		if (wrh_graph_run(g, iq.data())) { fprintf(stderr, "run failed at %d\n", b); rc = 1; }
	}
	stop.store(true);
	http1.join();
	http2.join();
	wrh_graph_destroy(g);
	printf("threads done\n");
	return rc;
}
