// dspblock.h -- the DspBlock / DspSource plugin surface, GPU edition.
//
// Same class names, public methods and protected hooks as WebRadio's src/dsp/dspblock.h
// (reference lines 47-140), so that the reference's Radio glue (src/radio.cxx) and web handlers
// compile against this header unchanged.  The graph semantics are the reference's:
//   * connect()/disconnect() hook blocks up (and start/stop the newcomer if already running);
//   * a source's start() cascades sample rate / channel count downstream, and every block's
//     input and output rates must be integer related (reference dspblock.cxx:106-151);
//   * run() sizes the block-owned output buffer, calls process(), then pushes the buffer to the
//     consumers depth-first in connection order (reference dspblock.cxx:169-212).
// Additions (public, used by the CUDA-backed blocks to batch work behind this per-block,
// synchronous interface): upstream()/downstream() graph access, a per-block run serial and a
// global topology serial.
#ifndef DSPBLOCK_H_
#define DSPBLOCK_H_

#define DSPBLOCK_PROFILE

#include <stdint.h>
#include <time.h>

#include <string>
#include <vector>

#define DEFAULT_SAMPLE_RATE 48000
#define DEFAULT_CHANNELS    2
#define DEFAULT_BLOCK_SIZE  16384

using namespace std;

typedef float sample_t;

class DspSource;

class DspBlock
{
public:
	friend class DspSource;

	DspBlock(const string &name = "<undefined>", const string &type = "DspBlock");
	virtual ~DspBlock();

	void connect(DspBlock *block);
	void disconnect(DspBlock *block);

	unsigned int inputSampleRate() const { return _inputSampleRate; }
	unsigned int outputSampleRate() const { return _outputSampleRate; }
	unsigned int inputChannels() const { return _inputChannels; }
	unsigned int outputChannels() const { return _outputChannels; }
	unsigned int decimation() const { return _decimation; }
	unsigned int interpolation() const { return _interpolation; }

#ifdef DSPBLOCK_PROFILE
	uint64_t nsPerFrameAll() const;
	// the reference divides by zero before the first block; this returns 0 instead
	uint64_t nsPerFrameOne() const { return _spent.framesIn ? _spent.ns / _spent.framesIn : 0; }
	uint64_t totalNanoseconds() const { return _spent.ns; }
	unsigned int totalIn() const { return (unsigned int)_spent.framesIn; }
	unsigned int totalOut() const { return (unsigned int)_spent.framesOut; }
#endif

	bool isRunning() const { return _isRunning; }
	const string &name() const { return _name; }
	const string &type() const { return _type; }

	// ---- additions (not in the reference) used by the GPU-backed blocks to batch work ----
	const vector<DspBlock*> &downstream() const { return _consumers; }
	DspBlock *upstream() const { return _producer; }
	// increments every time this block's run() is entered: lets a consumer tell whether the
	// producer buffer it is handed is a new block
	uint64_t runSerial() const { return _runSerial; }
	// increments on every connect()/disconnect() anywhere in the process
	static uint64_t topologySerial();
	// a consumer has started an asynchronous device copy of this block's output buffer
	// (wrhost::uploadFor): run() waits for it before it returns, stop() drops it before the buffer goes
	void noteUpload() { _uploadPending = _everUploaded = true; }

protected:
	// hooks a concrete block implements
	virtual bool init() { return false; }
	virtual void deinit() {}
	virtual bool process(const vector<sample_t> &inBuffer, vector<sample_t> &outBuffer)
	{
		(void)inBuffer; (void)outBuffer;
		return false;
	}

	// a block may adjust these from init()
	unsigned int _outputSampleRate;
	unsigned int _outputChannels;

private:
	// reachable only through DspSource (friend), exactly as in the reference
	bool start();
	void stop();
	bool run(const vector<sample_t> &inBuffer);
	void setSampleRate(unsigned int rate);
	void setChannels(unsigned int channels);

	// whole-number rate change in one direction, or failure (reference dspblock.cxx:119-130)
	bool negotiateRates();
	string label() const { return _type + ":" + _name; }

	// graph
	DspBlock *_producer;
	vector<DspBlock*> _consumers;
	vector<sample_t> _buffer;   // this block's output, handed to every consumer
	bool _isRunning;
	uint64_t _runSerial;
	bool _uploadPending, _everUploaded;

	// identity and negotiated format
	const string _name;
	const string _type;
	unsigned int _inputSampleRate;
	unsigned int _inputChannels;
	unsigned int _decimation;
	unsigned int _interpolation;

#ifdef DSPBLOCK_PROFILE
	struct Spent {
		uint64_t ns, framesIn, framesOut;
		Spent() : ns(0), framesIn(0), framesOut(0) {}
	} _spent;
#endif
};

class DspSource : public DspBlock
{
public:
	DspSource(const string &name = "<undefined>", const string &type = "DspSource");
	virtual ~DspSource();

	unsigned int blockSize() const { return _blockSize; }

	bool start() { return DspBlock::start(); }
	void stop() { DspBlock::stop(); }
	// a source has no input: it is run with a placeholder of blockSize() samples, which also
	// fixes the size of its output buffer (reference dspblock.h:134)
	bool run();
	void setSampleRate(unsigned int rate) { DspBlock::setSampleRate(rate); }
	void setChannels(unsigned int channels) { DspBlock::setChannels(channels); }
	void setBlockSize(unsigned int bytes);

private:
	unsigned int _blockSize;
	vector<sample_t> _placeholder;
};

#endif /* DSPBLOCK_H_ */
