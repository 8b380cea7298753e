// TEST STAND-IN for MTF's exception hierarchy (same class names and what()/type() surface), see opencv2/core/core.hpp here.
#pragma once
#include <stdexcept>
#include <string>
namespace mtf { namespace utils {
class Exception : public std::exception {
public:
	explicit Exception(const std::string &e) : error(e){}
	const char* what() const noexcept override{ return error.c_str(); }
	virtual const char* type() const noexcept = 0;
private:
	std::string error;
};
#define MTFB_STANDIN_EXC(name) class name : public Exception { public: explicit name(const std::string &e) : Exception(e){} \
	const char* type() const noexcept override{ return #name; } };
MTFB_STANDIN_EXC(InvalidTrackerState)
MTFB_STANDIN_EXC(FunctonNotImplemented)
MTFB_STANDIN_EXC(InvalidArgument)
MTFB_STANDIN_EXC(LogicError)
}}
