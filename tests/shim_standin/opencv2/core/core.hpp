// TEST STAND-IN (not OpenCV): the handful of cv::Mat members include/mtf_b200_tracker.h touches, so that the shim can be
// compiled and exercised in an image that has no OpenCV C++ headers.  Row-major, reference-counted like cv::Mat.
#pragma once
#include <cstddef>
#include <cstring>
#include <memory>
#include <vector>
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32FC1 5
#define CV_64FC1 6
namespace cv {
class Mat {
public:
	int rows = 0, cols = 0;
	size_t step = 0;
	unsigned char *data = nullptr;
	Mat(){}
	Mat(int r, int c, int t){ create(r, c, t); }
	Mat(int r, int c, int t, void *ext, size_t _step = 0) : rows(r), cols(c), step(_step ? _step : c*esz(t)), data((unsigned char*)ext), type_(t){}
	void create(int r, int c, int t){
		rows = r; cols = c; type_ = t; step = c*esz(t);
		store = std::make_shared<std::vector<unsigned char>>(r*step);
		data = store->data();
	}
	int type() const{ return type_; }
	bool empty() const{ return data == nullptr; }
	template<class T> T& at(int r, int c){ return *reinterpret_cast<T*>(data + r*step + c*sizeof(T)); }
	template<class T> const T& at(int r, int c) const{ return *reinterpret_cast<const T*>(data + r*step + c*sizeof(T)); }
	template<class T> T* ptr(int r = 0){ return reinterpret_cast<T*>(data + r*step); }
	template<class T> const T* ptr(int r = 0) const{ return reinterpret_cast<const T*>(data + r*step); }
	void copyTo(Mat &o) const{ o.create(rows, cols, type_); for(int r = 0; r < rows; ++r) std::memcpy(o.data + r*o.step, data + r*step, cols*esz(type_)); }
private:
	static size_t esz(int t){ return t == CV_64FC1 ? 8 : t == CV_8UC1 ? 1 : t == CV_8UC3 ? 3 : 4; }
	int type_ = 0;
	std::shared_ptr<std::vector<unsigned char>> store;
};
}
