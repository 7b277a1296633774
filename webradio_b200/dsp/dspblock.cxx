// dspblock.cxx -- graph plumbing of the DspBlock plugin surface (see dspblock.h).
// Behavioural contract: WebRadio src/dsp/dspblock.cxx (cited per function).
#include "dspblock.h"

#include <algorithm>
#include <atomic>

#define __STDC_FORMAT_MACROS
#include <inttypes.h>

#include "debug.h"

static std::atomic<uint64_t> g_topologySerial(1);

uint64_t DspBlock::topologySerial() { return g_topologySerial.load(); }

DspBlock::DspBlock(const string &name, const string &type) :
	_outputSampleRate(DEFAULT_SAMPLE_RATE), _outputChannels(DEFAULT_CHANNELS),
	_name(name), _type(type),
	_inputSampleRate(DEFAULT_SAMPLE_RATE), _inputChannels(DEFAULT_CHANNELS),
	_decimation(1), _interpolation(1),
#ifdef DSPBLOCK_PROFILE
	_totalNanoseconds(0), _totalIn(0), _totalOut(0),
#endif
	_isRunning(false), _runSerial(0), _producer(NULL)
{
}

DspBlock::~DspBlock()
{
	// reference dspblock.cxx:51-55
	if (_isRunning)
		stop();
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
// are refused.
void DspBlock::connect(DspBlock *block)
{
	if (_isRunning)
		block->start();
	if (std::find(_consumers.begin(), _consumers.end(), block) != _consumers.end()) {
		LOG_ERROR("%s:%s is already a consumer of %s:%s\n", block->type().c_str(), block->name().c_str(),
				type().c_str(), name().c_str());
		return;
	}
	_consumers.push_back(block);
	block->_producer = this;
	g_topologySerial++;
	LOG_DEBUG("%s:%s now feeds %s:%s\n", type().c_str(), name().c_str(), block->type().c_str(), block->name().c_str());
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
	LOG_DEBUG("%s:%s no longer feeds %s:%s\n", type().c_str(), name().c_str(), block->type().c_str(), block->name().c_str());
}

#ifdef DSPBLOCK_PROFILE
// reference dspblock.cxx:94-103: log this block's ns/frame, then add up the subtree
uint64_t DspBlock::nsPerFrameAll() const
{
	uint64_t sum = nsPerFrameOne();
	LOG_DEBUG("%s:%s %" PRIu64 " ns/frame\n", type().c_str(), name().c_str(), sum);
	for (size_t i = 0; i < _consumers.size(); i++)
		sum += _consumers[i]->nsPerFrameAll();
	return sum;
}
#endif

// reference dspblock.cxx:106-151
bool DspBlock::start()
{
	// a sink that never touches its output settings inherits its input's
	_outputSampleRate = _inputSampleRate;
	_outputChannels = _inputChannels;

	LOG_DEBUG("starting %s:%s\n", type().c_str(), name().c_str());
	if (!init()) {
		LOG_ERROR("%s:%s failed to initialise\n", type().c_str(), name().c_str());
		return false;
	}

	// whole-number rate change only, in one direction
	if (_outputSampleRate <= _inputSampleRate) {
		_interpolation = 1;
		_decimation = _outputSampleRate ? _inputSampleRate / _outputSampleRate : 0;
	} else {
		_decimation = 1;
		_interpolation = _inputSampleRate ? _outputSampleRate / _inputSampleRate : 0;
	}
	if (_decimation == 0 || _interpolation == 0 ||
			_inputSampleRate * _interpolation / _decimation != _outputSampleRate) {
		LOG_ERROR("Sample rates must be integer related\n");
		deinit();
		return false;
	}

#ifdef DSPBLOCK_PROFILE
	_totalNanoseconds = 0;
	_totalIn = _totalOut = 0;
#endif
	_isRunning = true;

	for (size_t i = 0; i < _consumers.size(); i++) {
		DspBlock *c = _consumers[i];
		c->setSampleRate(_outputSampleRate);
		c->setChannels(_outputChannels);
		if (!c->start()) {
			LOG_ERROR("downstream of %s:%s failed to start, stopping the pipeline\n", type().c_str(), name().c_str());
			stop();
			return false;
		}
	}
	return true;
}

// reference dspblock.cxx:153-167: consumers first, then this block; the buffer is released
void DspBlock::stop()
{
	for (size_t i = 0; i < _consumers.size(); i++)
		_consumers[i]->stop();
	if (_isRunning) {
		LOG_DEBUG("stopping %s:%s\n", type().c_str(), name().c_str());
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
		LOG_DEBUG("%s:%s output buffer -> %u frames x %u channels\n", type().c_str(), name().c_str(),
				outframes, _outputChannels);
		_buffer.resize(want);
	}

#ifdef DSPBLOCK_PROFILE
	timespec t0, t1;
	clock_gettime(CLOCK_PROCESS_CPUTIME_ID, &t0);
#endif
	if (!process(inBuffer, _buffer)) {
		LOG_ERROR("Pipeline failed at block %s:%s\n", type().c_str(), name().c_str());
		return false;
	}
#ifdef DSPBLOCK_PROFILE
	clock_gettime(CLOCK_PROCESS_CPUTIME_ID, &t1);
	_totalNanoseconds += (uint64_t)((int64_t)(t1.tv_sec - t0.tv_sec) * 1000000000LL + (int64_t)(t1.tv_nsec - t0.tv_nsec));
	_totalIn += inframes;
	_totalOut += outframes;
#endif

	for (size_t i = 0; i < _consumers.size(); i++)
		if (!_consumers[i]->run(_buffer))
			return false;
	return true;
}

// reference dspblock.cxx:214-231: ignored while running
void DspBlock::setSampleRate(unsigned int rate)
{
	if (_isRunning)
		return;
	_inputSampleRate = rate;
}

void DspBlock::setChannels(unsigned int channels)
{
	if (_isRunning)
		return;
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
