// tuner.h -- property bag of a tuner front-end (same surface as WebRadio's src/io/tuner.h:36-77).
// Hardware tuners are outside the hot path; the header exists so the reference's Radio glue and
// tuner drivers compile against this tree.
#ifndef TUNER_H_
#define TUNER_H_

#include <string>

#include "samplesource.h"

using namespace std;

#define DEFAULT_TUNER_SAMPLE_RATE 1200000
#define DEFAULT_TUNER_CHANNELS    2

class Tuner : public SampleSource
{
public:
	Tuner(const string &name = "<undefined>", const string &type = "Tuner") :
		SampleSource(name, type),
		_centreFrequency(100000000), _offsetPPM(0), _AGC(true), _gainDB(0) {}
	virtual ~Tuner() {}

	// device identification (these hide DspBlock::name() on purpose, as in the reference)
	const string &name() const { return _name; }
	const string &manufacturer() const { return _manufacturer; }
	const string &product() const { return _product; }
	const string &serial() const { return _serial; }

	unsigned int centreFrequency() const { return _centreFrequency; }
	int offsetPPM() const { return _offsetPPM; }
	bool AGC() const { return _AGC; }
	virtual float gainDB() const { return _gainDB; }

	virtual void setCentreFrequency(unsigned int hz) { (void)hz; }
	virtual void setOffsetPPM(int ppm) { (void)ppm; }
	virtual void setAGC(bool agc) { (void)agc; }
	virtual void setGainDB(float gain) { (void)gain; }

protected:
	string _name, _manufacturer, _product, _serial;
	unsigned int _centreFrequency;
	int _offsetPPM;
	bool _AGC;
	float _gainDB;
};

typedef Tuner *(*TunerFactory)(const string&);

#endif /* TUNER_H_ */
