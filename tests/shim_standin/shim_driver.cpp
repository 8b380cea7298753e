// TEST DRIVER: compiles include/mtf_b200_tracker.h against the stand-in MTF / OpenCV headers and drives it the way runMTF
// (Examples/cpp/runMTF.cc:154,692-704) and GridTracker (SM/src/GridTracker.cc:232-264) drive a TrackerBase.
// usage: shim_driver <frames.bin: int32 n,h,w + n*h*w float32> <corners.bin: int32 P + P*8 float64> sm am ssm res -> prints corners
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "mtf_b200_tracker.h"

int main(int argc, char **argv){
	if(argc < 7){ fprintf(stderr, "usage\n"); return 2; }
	FILE *f = fopen(argv[1], "rb"); int n, h, w;
	if(!f || fread(&n, 4, 1, f) != 1 || fread(&h, 4, 1, f) != 1 || fread(&w, 4, 1, f) != 1) return 2;
	std::vector<float> frames((size_t)n*h*w);
	if(fread(frames.data(), 4, frames.size(), f) != frames.size()) return 2;
	fclose(f);
	f = fopen(argv[2], "rb"); int P;
	if(!f || fread(&P, 4, 1, f) != 1) return 2;
	std::vector<double> corners((size_t)P * 8);
	if(fread(corners.data(), 8, corners.size(), f) != corners.size()) return 2;
	fclose(f);
	const int res = atoi(argv[6]);
	try{
		// the application keeps ONE image buffer and overwrites it in place every frame (TrackerBase.h:21-26)
		cv::Mat img(h, w, CV_32FC1);
		auto load = [&](int t){ for(int r = 0; r < h; ++r) memcpy(img.ptr<float>(r), &frames[((size_t)t*h + r)*w], w * 4); };
		// (1) single trackers, one per patch, as runMTF would create them
		std::vector<mtf::TrackerBase*> single;
		for(int i = 0; i < P; ++i) single.push_back(new mtf::b200::Tracker(argv[3], argv[4], argv[5], res, res));
		// (2) the same patches as members of one batch, as a GridTracker would hold them
		std::vector<mtf::TrackerBase*> members = mtf::b200::makeBatchMembers(argv[3], argv[4], argv[5], P, res, res);
		load(0);
		for(int i = 0; i < P; ++i){
			cv::Mat c(2, 4, CV_64FC1);
			for(int k = 0; k < 8; ++k) c.at<double>(k / 4, k % 4) = corners[8 * (size_t)i + k];
			single[i]->initialize(img, c);
			members[i]->setImage(img); members[i]->initialize(c);
			// GridTracker::resetTrackers reads the region right after initialising a member (GridTracker.cc:381-387): it must
			// be the supplied one even though the batch is only (re)initialised when its last member has arrived
			const cv::Mat &r = members[i]->getRegion();
			for(int k = 0; k < 8; ++k) if(r.at<double>(k / 4, k % 4) != corners[8 * (size_t)i + k]){ printf("STALE %d\n", i); break; }
		}
		for(int t = 1; t < n; ++t){
			load(t);
			for(int i = 0; i < P; ++i){ single[i]->update(); members[i]->update(); }
		}
		for(int i = 0; i < P; ++i){
			const cv::Mat &a = single[i]->getRegion(), &b = members[i]->getRegion();
			for(int k = 0; k < 8; ++k) printf("%.17g %.17g\n", a.at<double>(k / 4, k % 4), b.at<double>(k / 4, k % 4));
		}
		// (3) raw uint8 frames with the pre-processing on the device (Batch::useRawInput), as an application would feed a
		// tracker whose inputType() is CV_8UC1 through a "none" pre-processor
		if(argc > 7 && !strcmp(argv[7], "raw")){
			mtf::b200::Tracker raw(argv[3], argv[4], argv[5], res, res);
			raw.getBatch().useRawInput(1);
			if(raw.inputType() != CV_8UC1){ printf("BADTYPE\n"); }
			cv::Mat img8(h, w, CV_8UC1);
			auto load8 = [&](int t){ for(int r = 0; r < h; ++r) for(int c = 0; c < w; ++c){
				float v = frames[((size_t)t*h + r)*w + c]; v = v < 0 ? 0 : (v > 255 ? 255 : v); img8.ptr<unsigned char>(r)[c] = (unsigned char)(v + 0.5f); } };
			load8(0);
			cv::Mat c(2, 4, CV_64FC1);
			for(int k = 0; k < 8; ++k) c.at<double>(k / 4, k % 4) = corners[k];
			raw.initialize(img8, c);
			for(int t = 1; t < n; ++t){ load8(t); raw.update(); }
			const cv::Mat &a = raw.getRegion();
			for(int k = 0; k < 8; ++k) printf("RAW %.17g\n", a.at<double>(k / 4, k % 4));
		}
		// (4) the particle filter as a TrackerBase (mtf::b200::PFTracker), device generator with a fixed seed
		if(argc > 7 && !strcmp(argv[7], "raw") && !strcmp(argv[5], "8")){
			mtfb_pf_params pf;
			mtfb_pf_default_params(&pf);
			pf.n_particles = 300; pf.ssm_sigma[0] = 0.5; pf.ssm_sigma[1] = 0.2; pf.seed = 77; pf.mean_type = MTFB_PF_MEAN_SSM;
			mtf::b200::PFTracker pft(argv[4], argv[5], res, res, pf);
			load(0);
			cv::Mat c(2, 4, CV_64FC1);
			for(int k = 0; k < 8; ++k) c.at<double>(k / 4, k % 4) = corners[k];
			pft.initialize(img, c);
			for(int t = 1; t < n; ++t){ load(t); pft.update(); }
			const cv::Mat &a = pft.getRegion();
			for(int k = 0; k < 8; ++k) printf("PF %.17g\n", a.at<double>(k / 4, k % 4));
			try{ mtf::b200::Tracker bad("pf", argv[4], argv[5], res, res); printf("NOEXCPF\n"); }
			catch(const mtf::utils::Exception &e){ printf("EXCPF %s\n", e.type()); }
		}
		// (5) the step after the cells in GridTracker::update: the warp of the region from the members' centroids, on the device
		// through the batch the members share (Batch::gridEstimate) and from host points (Batch::estimateWarpFromPts)
		{
			std::shared_ptr<mtf::b200::Batch> gb(new mtf::b200::Batch(mtf::b200::makeParams(argv[3], argv[4], argv[5], P, res, res)));
			gb->gridEnable();
			load(0); gb->setImage(img); gb->initialize(corners.data());
			std::vector<float> prev(2 * (size_t)P), curr(2 * (size_t)P);
			for(int i = 0; i < P; ++i){ const double *r = gb->region(i);
				prev[2 * i] = (float)((r[0] + r[1] + r[2] + r[3]) / 4.0); prev[2 * i + 1] = (float)((r[4] + r[5] + r[6] + r[7]) / 4.0); }
			load(1); gb->setImage(img); gb->update();
			for(int i = 0; i < P; ++i){ const double *r = gb->region(i);
				curr[2 * i] = (float)((r[0] + r[1] + r[2] + r[3]) / 4.0); curr[2 * i + 1] = (float)((r[4] + r[5] + r[6] + r[7]) / 4.0); }
			const mtfb_est_params ep = mtf::b200::Batch::estParams(MTFB_EST_RANSAC, 10.0, 4, true, 2000, 300, 0.995, 10, 99);
			double su_dev[8], su_host[8]; std::vector<unsigned char> m_dev(P), m_host(P);
			const bool ok_dev = gb->gridEstimate(MTFB_SSM_HOMOGRAPHY, su_dev, m_dev.data(), ep);
			const bool ok_host = gb->estimateWarpFromPts(MTFB_SSM_HOMOGRAPHY, su_host, m_host.data(), prev.data(), curr.data(), P, ep);
			printf("EST %d %d\n", (int)ok_dev, (int)ok_host);
			for(int k = 0; k < 8; ++k) printf("ESTSU %.17g %.17g\n", su_dev[k], su_host[k]);
			for(int i = 0; i < P; ++i) if(m_dev[i] != m_host[i]){ printf("ESTMASK %d\n", i); }
			for(int i = 0; i < P; ++i) printf("ESTPT %.9g %.9g %.9g %.9g\n", prev[2 * i], prev[2 * i + 1], curr[2 * i], curr[2 * i + 1]);
		}
		// (5) mtf::b200::GridTracker: a 4 x 4 grid of 12 x 12 ESM + NCC + Affine cells under a Homography, driven like any TrackerBase
		if(argc > 7 && !strcmp(argv[7], "raw")){
			const mtfb_est_params ep = mtf::b200::Batch::estParams(MTFB_EST_RANSAC, 2.0, 4, true, 2000, 300, 0.995, 10, 5);
			mtf::b200::GridTracker grid("esm", "ncc", "6", 4, 4, 20, 22, 12, 12, "8", ep);
			cv::Mat c(2, 4, CV_64FC1);
			const double reg[8] = { 90.0, 300, 305, 85, 80, 84, 290, 296 };
			for(int k = 0; k < 8; ++k) c.at<double>(k / 4, k % 4) = reg[k];
			load(0); grid.initialize(img, c);
			for(int t = 1; t < n; ++t){
				load(t); grid.update(img);
				const cv::Mat &r = grid.getRegion();
				for(int k = 0; k < 8; ++k) printf("GRID %d %.17g\n", t, r.at<double>(k / 4, k % 4));
			}
		}
		// error contract: a bad corner matrix is an InvalidArgument exception, not a crash
		try{ cv::Mat bad(3, 4, CV_64FC1); single[0]->initialize(bad); printf("NOEXC\n"); }
		catch(const mtf::utils::Exception &e){ printf("EXC %s\n", e.type()); }
	} catch(const mtf::utils::Exception &e){
		fprintf(stderr, "%s: %s\n", e.type(), e.what()); return 1;
	}
	return 0;
}
