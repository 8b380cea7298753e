// TEST STAND-IN for mtf::TrackerBase: the same virtual interface (names, signatures, convenience overloads), nothing else.
#pragma once
#include "opencv2/core/core.hpp"
#include "mtf/Utilities/excpUtils.h"
#include <string>
namespace mtf {
class TrackerBase {
protected:
	cv::Mat cv_corners_mat;
public:
	std::string name;
	TrackerBase(){}
	virtual ~TrackerBase(){}
	virtual void setImage(const cv::Mat &img) = 0;
	virtual void initialize(const cv::Mat &corners) = 0;
	virtual void update() = 0;
	virtual void setRegion(const cv::Mat &corners){ (void)corners; throw mtf::utils::FunctonNotImplemented(name + " :: setRegion"); }
	virtual void initialize(const cv::Mat &img, const cv::Mat &corners){ setImage(img); initialize(corners); }
	virtual void update(const cv::Mat &img){ setImage(img); update(); }
	virtual void setRegion(const cv::Mat &img, const cv::Mat &corners){ setImage(img); setRegion(corners); }
	virtual const cv::Mat& getRegion(){ return cv_corners_mat; }
	virtual int inputType() const = 0;
};
}
