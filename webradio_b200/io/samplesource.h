// samplesource.h -- base of blocks that originate a pipeline (same surface as WebRadio's
// src/io/samplesource.h:33-60).
#ifndef SAMPLESOURCE_H_
#define SAMPLESOURCE_H_

#include <string>
#include <vector>

#include "dspblock.h"

using namespace std;

class SampleSource : public DspSource
{
public:
	SampleSource(const string &name = "<undefined>", const string &type = "SampleSource") :
		DspSource(name, type) {}
	virtual ~SampleSource() {}

	const string &subdevice() const { return _subdevice; }
	const vector<string> &subdevices() const { return _subdevices; }
	void setSubdevice(const string &subdevice)
	{
		if (!isRunning())
			_subdevice = subdevice;
	}

protected:
	vector<string> _subdevices; // filled in by the concrete source's constructor

private:
	string _subdevice;
};

#endif /* SAMPLESOURCE_H_ */
