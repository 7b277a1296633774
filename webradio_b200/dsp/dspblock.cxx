// dspblock.cxx -- graph plumbing of the DspBlock plugin surface (see dspblock.h).
// Behavioural contract: WebRadio src/dsp/dspblock.cxx (cited per function).
#include "dspblock.h"

#include <algorithm>
#include <atomic>

#define __STDC_FORMAT_MACROS
#include <inttypes.h>

#include "debug.h"
#include "gpubank.h"

static std::atomic<uint64_t> g_topologySerial(1);

uint64_t DspBlock::topologySerial() { return g_topologySerial.load(); }

DspBlock::DspBlock(const string &name, const string &type) :
	_outputSampleRate(DEFAULT_SAMPLE_RATE), _outputChannels(DEFAULT_CHANNELS),
	_producer(NULL), _isRunning(false), _runSerial(0), _uploadPending(false), _everUploaded(false),
	_name(name), _type(type),
	_inputSampleRate(DEFAULT_SAMPLE_RATE), _inputChannels(DEFAULT_CHANNELS),
	_decimation(1), _interpolation(1)
{
}

DspBlock::~DspBlock()
{
	// reference dspblock.cxx:51-55
	if (_isRunning)
		stop();
	if (_everUploaded)
		wrhost::forget(this);
	// Blocks may be deleted in either order (the reference's destructor never looks at its
	// consumers unless it is running).  The upstream() link added here must not change that: a
	// consumer that goes away takes itself off its producer's list, so the loop below only ever
	// walks over live blocks.
	if (_producer) {
		vector<DspBlock*> &siblings = _producer->_consumers;
		siblings.erase(std::remove(siblings.begin(), siblings.end(), this), siblings.end());
		g_topologySerial++;
	}
	for (size_t i = 0; i < _consumers.size(); i++)
		if (_consumers[i]->_producer == this)
			_consumers[i]->_producer = NULL;
}

// reference dspblock.cxx:57-76: a block joining a running pipeline is started on the spot
// (without a rate cascade -- the newcomer keeps whatever rates it was last given); duplicates
// are refused -- after that start, as in the reference.
void DspBlock::connect(DspBlock *block)
{
	if (_isRunning)
		block->start();
	const bool known = std::find(_consumers.begin(), _consumers.end(), block) != _consumers.end();
	if (known) {
		LOG_ERROR("%s is already a consumer of %s\n", block->label().c_str(), label().c_str());
		return;
	}
	block->_producer = this;
	_consumers.push_back(block);
	g_topologySerial++;
	LOG_DEBUG("%s now feeds %s\n", label().c_str(), block->label().c_str());
}

// reference dspblock.cxx:78-91
void DspBlock::disconnect(DspBlock *block)
{
	if (_isRunning)
		block->stop();
	_consumers.erase(std::remove(_consumers.begin(), _consumers.end(), block), _consumers.end());
	if (block->_producer == this)
		block->_producer = NULL;
	g_topologySerial++;
	LOG_DEBUG("%s no longer feeds %s\n", label().c_str(), block->label().c_str());
}

#ifdef DSPBLOCK_PROFILE
// reference dspblock.cxx:94-103: log this block's ns/frame, then add up the subtree
uint64_t DspBlock::nsPerFrameAll() const
{
	uint64_t sum = nsPerFrameOne();
	LOG_DEBUG("%s %" PRIu64 " ns/frame\n", label().c_str(), sum);
	for (size_t i = 0; i < _consumers.size(); i++)
		sum += _consumers[i]->nsPerFrameAll();
	return sum;
}

// The reference reads CLOCK_PROCESS_CPUTIME_ID around every process() (dspblock.cxx:186-204): a
// system call each, 500 of them per tuner block with 64 receivers -- a fifth of a millisecond spent
// on measuring -- and CPU time says nothing about a block that waits for the GPU.  The monotonic
// clock (a vDSO read, tens of nanoseconds) keeps the counters and their getters.
static inline uint64_t cpuTimeNs()
{
	timespec t;
	clock_gettime(CLOCK_MONOTONIC, &t);
	return (uint64_t)t.tv_sec * 1000000000ull + (uint64_t)t.tv_nsec;
}
#endif

bool DspBlock::negotiateRates()
{
	const unsigned int in = _inputSampleRate, out = _outputSampleRate;
	_decimation = _interpolation = 1;
	if (in == 0 || out == 0)
		return false;
	if (out <= in)
		_decimation = in / out;
	else
		_interpolation = out / in;
	return in * _interpolation / _decimation == out;
}

// reference dspblock.cxx:106-151
bool DspBlock::start()
{
	// a sink that never touches its output settings inherits its input's
	_outputSampleRate = _inputSampleRate;
	_outputChannels = _inputChannels;

	LOG_DEBUG("starting %s\n", label().c_str());
	if (!init()) {
		LOG_ERROR("%s failed to initialise\n", label().c_str());
		return false;
	}
	if (!negotiateRates()) {
		LOG_ERROR("Sample rates must be integer related\n");
		deinit();
		return false;
	}

#ifdef DSPBLOCK_PROFILE
	_spent = Spent();
#endif
	_isRunning = true;

	// hand the format downstream; one consumer that cannot start takes the pipeline down
	for (size_t i = 0; i < _consumers.size(); i++) {
		DspBlock *c = _consumers[i];
		c->setSampleRate(_outputSampleRate);
		c->setChannels(_outputChannels);
		if (c->start())
			continue;
		LOG_ERROR("downstream of %s failed to start, stopping the pipeline\n", label().c_str());
		stop();
		return false;
	}
	return true;
}

// reference dspblock.cxx:153-167: consumers first, then this block; the buffer is released
void DspBlock::stop()
{
	for (size_t i = 0; i < _consumers.size(); i++)
		_consumers[i]->stop();
	if (_everUploaded) {
		// this block's buffers are page-locked in place for the device copies: undo that before
		// deinit() or the swap below frees them
		wrhost::forget(this);
		_uploadPending = _everUploaded = false;
	}
	if (_isRunning) {
		LOG_DEBUG("stopping %s\n", label().c_str());
		_isRunning = false;
		deinit();
	}
	vector<sample_t>().swap(_buffer);
}

// reference dspblock.cxx:169-212
bool DspBlock::run(const vector<sample_t> &inBuffer)
{
	if (!_isRunning) {
		LOG_ERROR("Pipeline not started\n");
		return false;
	}
	_runSerial++;

	// truncating frame arithmetic, as the reference (dspblock.cxx:177-178)
	const unsigned int inframes = _inputChannels ? (unsigned int)(inBuffer.size() / _inputChannels) : 0;
	const unsigned int outframes = inframes * _interpolation / _decimation;
	const size_t want = (size_t)outframes * _outputChannels;
	if (_buffer.size() != want) {
		LOG_DEBUG("%s output buffer -> %u frames x %u channels\n", label().c_str(), outframes, _outputChannels);
		_buffer.resize(want);
	}

#ifdef DSPBLOCK_PROFILE
	const uint64_t t0 = cpuTimeNs();
#endif
	if (!process(inBuffer, _buffer)) {
		LOG_ERROR("Pipeline failed at block %s\n", label().c_str());
		return false;
	}
#ifdef DSPBLOCK_PROFILE
	_spent.ns += cpuTimeNs() - t0;
	_spent.framesIn += inframes;
	_spent.framesOut += outframes;
#endif

	bool ok = true;
	for (size_t i = 0; ok && i < _consumers.size(); i++)
		ok = _consumers[i]->run(_buffer);
	if (_uploadPending) {
		// a consumer copies this buffer to its device asynchronously: it is ours again only now
		wrhost::blockDone(this);
		_uploadPending = false;
	}
	return ok;
}

// reference dspblock.cxx:214-231: ignored while running
void DspBlock::setSampleRate(unsigned int rate)
{
	if (!_isRunning)
		_inputSampleRate = rate;
}

void DspBlock::setChannels(unsigned int channels)
{
	if (!_isRunning)
		_inputChannels = channels;
}

DspSource::DspSource(const string &name, const string &type) :
	DspBlock(name, type), _blockSize(DEFAULT_BLOCK_SIZE)
{
}

DspSource::~DspSource()
{
}

bool DspSource::run()
{
	if (_placeholder.size() != _blockSize)
		_placeholder.assign(_blockSize, 0.0f);
	return DspBlock::run(_placeholder);
}

// reference dspblock.cxx:242-249: ignored while running
void DspSource::setBlockSize(unsigned int size)
{
	if (isRunning())
		return;
	_blockSize = size;
}
