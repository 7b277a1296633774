// tests/harness/host_scenario.cxx -- TEST DRIVER: the drop-in blocks' host logic under AddressSanitizer
// and UBSan (make asan-check; stand-in back-end of mock_capi.cxx, so no GPU is needed).  Two live
// pipelines with fused, strict and mixed chains and a spectrum sink go through detach / attach /
// double attach / setters / restart / detach-everything and are torn down consumers-first and
// producers-first.  What it found when it was written: ~DspBlock walked over consumers that had been
// deleted before their producer (heap-use-after-free; the reference allows either order).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
extern "C" {
void *wrh_graph_create(unsigned fs, unsigned block_frames);
int wrh_graph_add_receiver(void *h, int if_hz, unsigned ch_passband, unsigned ch_rate, unsigned ch_decim, int mode,
		unsigned au_passband, unsigned au_rate, unsigned au_decim, unsigned capture_mask);
int wrh_graph_add_spectrum(void *h, unsigned fft_size);
int wrh_graph_start(void *h);
int wrh_graph_run(void *h, const float *iq);
long wrh_graph_get(void *h, int rx, int stage, float *out, long cap);
int wrh_graph_detach(void *h, int rx);
int wrh_graph_attach(void *h, int rx);
int wrh_graph_restart(void *h);
int wrh_graph_set_if(void *h, int rx, int hz);
int wrh_graph_set_mode(void *h, int rx, const char *mode);
int wrh_graph_set_passband(void *h, int rx, int which, unsigned hz);
int wrh_graph_spectrum(void *h, float *db);
void wrh_graph_destroy(void *h);
}
extern "C" void wrh_set_quiet(int);
int main()
{
	wrh_set_quiet(0);
	const unsigned fs = 2400000, F = 20000;
	std::vector<float> iq(2 * F);
	unsigned s = 12345;
	double sum = 0;
	for (int round = 0; round < 3; round++) {
		void *g = wrh_graph_create(fs, F);
		void *g2 = wrh_graph_create(fs, F);
		for (int i = 0; i < 6; i++)
			wrh_graph_add_receiver(g, 10000 * i - 7, 80000, 240000, 0, i % 4, 8000, 48000, 0, i == 3 ? 0xF : (i == 4 ? 0x9 : 0x8));
		for (int i = 0; i < 2; i++)
			wrh_graph_add_receiver(g2, 5000 * i, 12500, 0, 50, 1, 3000, 0, 1, 0x8);
		wrh_graph_add_spectrum(g, 512);
		{ int a = wrh_graph_start(g), c = wrh_graph_start(g2); if (a || c) { fprintf(stderr, "start failed %d %d\n", a, c); return 1; } }
		std::vector<float> audio(4096), db(512);
		for (int b = 0; b < 14; b++) {
			for (size_t k = 0; k < iq.size(); k++) { s = s * 1664525u + 1013904223u; iq[k] = ((float)(s >> 24) - 128.0f) / 128.0f; }
			if (b == 2) { wrh_graph_detach(g, 1); wrh_graph_detach(g, 3); }
			if (b == 3) { wrh_graph_set_if(g, 0, -99999); wrh_graph_set_mode(g, 2, "FM"); wrh_graph_set_passband(g, 5, 0, 200000); }
			if (b == 4) { wrh_graph_attach(g, 1); }
			if (b == 7 && round == 0) {
				// a receiver created on a live radio (POST /receivers in the reference's web UI): it starts
				// with the default 48 kHz input rate -- meaningless output, but it must not corrupt anything
				wrh_graph_add_receiver(g, 4242, 80000, 240000, 0, 1, 8000, 48000, 0, 0x8);
				wrh_graph_add_receiver(g, -4242, 80000, 240000, 0, 0, 8000, 48000, 0, 0xF);
			}
			if (b == 5) { wrh_graph_detach(g, 0); wrh_graph_detach(g, 2); wrh_graph_detach(g, 4); wrh_graph_detach(g, 5); wrh_graph_detach(g, 1); }
			if (b == 6) { wrh_graph_attach(g, 3); wrh_graph_attach(g, 5); }
			if (b == 8) { wrh_graph_restart(g); wrh_graph_restart(g2); }
			if (b == 10) { for (int i = 0; i < 6; i++) wrh_graph_attach(g, i); }
			if (b == 12) { wrh_graph_detach(g2, 0); }
			if (wrh_graph_run(g, iq.data()) || wrh_graph_run(g2, iq.data())) { fprintf(stderr, "run failed at %d\n", b); return 1; }
			for (int i = 0; i < 6; i++) {
				long n = wrh_graph_get(g, i, 3, audio.data(), 4096);
				for (long k = 0; k < n && k < 4096; k++) sum += audio[k];
			}
			wrh_graph_spectrum(g, db.data());
			sum += std::isfinite(db[7]) ? db[7] : 0;
		}
		if (round == 1) { wrh_graph_destroy(g2); wrh_graph_destroy(g); }   // either order
		else { wrh_graph_destroy(g); wrh_graph_destroy(g2); }
	}
	printf("scenario done, checksum %.6f\n", sum);
	return 0;
}
