// downconverter.cxx -- DownConverter block over libwebradio_b200.
// Behavioural contract: WebRadio src/dsp/downconverter.cxx:41-114.
#include "downconverter.h"

#include "debug.h"
#include "gpubank.h"
#include "webradio_b200.h"

DownConverter::DownConverter(const string &name) :
	DspBlock(name, "DownConverter"),
	filter(new LowPass(name)),
	_if(0),
	phase(0), phaseStep(0),   // the phase is initialised here only, never on (re)start
	bank(NULL), bankSlot(-1), plannedTopology(0), stage(NULL)
{
	// The 65536-entry sine table of the reference's constructor (downconverter.cxx:49-51) is
	// built once per device by libwebradio_b200 (wr_build_sintable) and kept in HBM.
}

DownConverter::~DownConverter()
{
	if (bank)
		wrhost::release(bank, bankSlot);
	wrhost::forget(this);   // the device copy of a chain that had no producer
	if (stage)
		wr_stage_destroy(stage);
	delete filter;
}

// reference downconverter.cxx:59-67: the step only follows the IF while running; init()
// recomputes it from _if on the next start.  Called from HTTP threads; the new step is picked
// up at the next block boundary.
void DownConverter::setIF(int hz)
{
	_if = hz;
	if (isRunning())
		phaseStep = wr_phase_step(hz, inputSampleRate());
}

// reference downconverter.cxx:69-84
bool DownConverter::init()
{
	if (inputChannels() != 2) {
		LOG_ERROR("Expect IQ input\n");
		return false;
	}
	_outputSampleRate = inputSampleRate();
	_outputChannels = inputChannels();
	phaseStep = wr_phase_step(_if, inputSampleRate());
	LOG_DEBUG("phaseStep = %d for %d Hz\n", (int)phaseStep, (int)_if);
	plannedTopology = 0; // re-plan on the first block
	return true;
}

void DownConverter::deinit()
{
	if (bank) {
		phase = bank->phaseOf(bankSlot);
		wrhost::release(bank, bankSlot);
		bank = NULL;
		bankSlot = -1;
	}
}

bool DownConverter::process(const vector<sample_t> &inBuffer, vector<sample_t> &outBuffer)
{
	const unsigned int nframes = (unsigned int)(inBuffer.size() / inputChannels());

	// (re)plan when the graph changed: is this chain fusable, and with whom?
	const uint64_t topo = DspBlock::topologySerial();
	if (topo != plannedTopology) {
		plannedTopology = topo;
		int slot = -1;
		wrhost::FusedBank *now = wrhost::planFor(this, &slot);
		if (now != bank || slot != bankSlot) {
			if (bank) {
				phase = bank->phaseOf(bankSlot); // carry the NCO phase over to the new back-end
				wrhost::release(bank, bankSlot);
			}
			bank = now;
			bankSlot = slot;
		}
	}

	if (bank) {
		// fused path: the first receiver of the bank to see this producer block runs the
		// kernels for all of them; the mixed IQ is never materialised (outBuffer keeps its size
		// but not meaningful contents -- nothing but the fused chain consumes it)
		// (a chain without a producer is its own clock: every call is a new block)
		DspBlock *src = upstream();
		return bank->ensureProcessed(src ? src->runSerial() : runSerial(), inBuffer.data(), nframes, this);
	}

	// strict path: this block alone, one kernel (reference downconverter.cxx:91-114)
	if (!stage) {
		stage = wr_stage_create(wrhost::deviceFor(upstream()));
		if (!stage) {
			LOG_ERROR("DownConverter: %s\n", wr_last_error());
			return false;
		}
	}
	if (wr_stage_mix(stage, NULL, &phase, phaseStep, inBuffer.data(), nframes, outBuffer.data()) != WR_OK) {
		LOG_ERROR("DownConverter: %s\n", wr_last_error());
		return false;
	}
	return true;
}
