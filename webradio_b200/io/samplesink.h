// samplesink.h -- base of blocks that terminate a pipeline (same surface as WebRadio's
// src/io/samplesink.h:33-58: a list of sub-devices and the selected one).
#ifndef SAMPLESINK_H_
#define SAMPLESINK_H_

#include <string>
#include <vector>

#include "dspblock.h"

using namespace std;

class SampleSink : public DspBlock
{
public:
	SampleSink(const string &name = "<undefined>", const string &type = "SampleSink") :
		DspBlock(name, type) {}
	virtual ~SampleSink() {}

	const string &subdevice() const { return _subdevice; }
	const vector<string> &subdevices() const { return _subdevices; }
	void setSubdevice(const string &subdevice)
	{
		if (!isRunning())
			_subdevice = subdevice;
	}

protected:
	vector<string> _subdevices;

private:
	string _subdevice;
};

#endif /* SAMPLESINK_H_ */
