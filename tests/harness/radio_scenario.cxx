// tests/harness/radio_scenario.cxx -- TEST DRIVER: the reference's UNMODIFIED src/radio.cxx (FrontEnd / Receiver
// life cycle, Radio::run) on the drop-in blocks under AddressSanitizer + UBSan, stand-in back-end (make
// asan-check, only where the reference tree is mounted).  Two create / run / retune / destroy rounds.
#include <cstdio>
#include <vector>
extern "C" {
void *wrr_create(unsigned fs, unsigned block_frames, unsigned fft_size);
int wrr_add_receiver(void *h, int if_hz, const char *mode);
int wrr_start(void *h);
int wrr_run(void *h, const float *iq);
int wrr_retune(void *h, int rx, int if_hz, const char *mode, unsigned chan_passband);
long wrr_audio(void *h, int rx, float *out, long cap);
int wrr_spectrum(void *h, float *db);
void wrr_destroy(void *h);
}
int main()
{
	const unsigned fs = 2400000, F = 102400;
	for (int round = 0; round < 2; round++) {
		void *rig = wrr_create(fs, F, 512);
		const char *modes[] = { "AM", "USB", "FM", "LSB" };
		for (int i = 0; i < 4; i++)
			if (wrr_add_receiver(rig, 100000 * i - 150000, modes[i]) < 0) return 1;
		if (wrr_start(rig)) { fprintf(stderr, "start failed\n"); return 1; }
		std::vector<float> iq(2 * F), audio(4096), db(512);
		unsigned s = 7;
		double sum = 0;
		for (int b = 0; b < 4; b++) {
			for (size_t k = 0; k < iq.size(); k++) { s = s * 1664525u + 1013904223u; iq[k] = ((float)(s >> 24) - 128.0f) / 128.0f; }
			if (b == 2) wrr_retune(rig, 1, -77777, "LSB", 100000);
			if (wrr_run(rig, iq.data())) { fprintf(stderr, "run failed\n"); return 1; }
			for (int i = 0; i < 4; i++) { long n = wrr_audio(rig, i, audio.data(), 4096); for (long k = 0; k < n && k < 4096; k++) sum += audio[k]; }
			wrr_spectrum(rig, db.data());
		}
		wrr_destroy(rig);
		printf("round %d checksum %.6f\n", round, sum);
	}
	printf("radio glue done\n");
	return 0;
}
