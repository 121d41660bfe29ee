// DsrcModuleGpu -- wrap::DsrcModule (reference include/dsrc/DsrcModule.h:22-40, src/DsrcModule.cpp:34-89) with the GPU operators
// behind it: the public library entry point of DSRC (the class the reference's boost::python module and example programs use),
// same Configurable setters, same Compress/Decompress(file, file), errors as DsrcException with the operator's message.
#pragma once
#include "../include/dsrc/Configurable.h"   // resolved against the reference tree: -I<reference>/src
#include "DsrcOperatorGpu.h"

namespace dsrc { namespace wrap {

class DsrcModuleGpu : public Configurable
{
public:
	void Compress(const std::string& inputFilename_, const std::string& outputFilename_)
	{
		comp::DsrcCompressorGpu op;
		Run(op, inputFilename_, outputFilename_);
	}

	void Decompress(const std::string& inputFilename_, const std::string& outputFilename_)
	{
		comp::DsrcDecompressorGpu op;
		Run(op, inputFilename_, outputFilename_);
	}

private:
	void Run(comp::IDsrcOperator& op_, const std::string& inputFilename_, const std::string& outputFilename_)
	{
		comp::InputParameters params = *(const comp::InputParameters*)GetInputParameters();			// DsrcModule.cpp:52
		params.inputFilename = inputFilename_;
		params.outputFilename = outputFilename_;
		if (!op_.Process(params))
			throw DsrcException(op_.GetError());
	}

	using Configurable::IsColorSpace;
	using Configurable::SetColorSpace;
	using Configurable::IsPlusRepetition;
	using Configurable::SetPlusRepetition;
};

} }
