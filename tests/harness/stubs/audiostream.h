// TEST-ONLY stand-in for WebRadio's src/web/audiostream.h, which pulls in libmicrohttpd and LAME
// (neither is installed here, and the web/MP3 layer is outside the hot path).  It declares just
// what src/radio.{h,cxx} use of AudioStreamManager: a SampleSink constructed from a name.  This
// one keeps the last audio block so the drop-in test can read what the receiver produced.
#ifndef AUDIOSTREAM_H_
#define AUDIOSTREAM_H_

#include <string>
#include <vector>

#include "samplesink.h"

class AudioStreamManager : public SampleSink
{
public:
	AudioStreamManager(const string &name = "<undefined>") : SampleSink(name, "AudioStreamManager"), blocks(0) {}
	vector<float> last;
	unsigned long blocks;
	// the reference's manager returns at once while no HTTP client is connected (audiostream.cxx:65-73):
	// capture(false) makes the stand-in do the same (bench.py's timed region)
	static bool &capture() { static bool on = true; return on; }
private:
	bool init() { return true; }
	void deinit() {}
	bool process(const vector<sample_t> &in, vector<sample_t> &out)
	{
		(void)out;
		if (capture())
			last.assign(in.begin(), in.end());
		blocks++;
		return true;
	}
};

#endif
