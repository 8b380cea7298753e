/*
 * mtf_b200_tracker.h -- header-only C++ shim that plugs libmtf_b200.so into MTF.
 *
 * Compiled INSIDE the MTF tree (it includes MTF's own headers); it contains no arithmetic, only the mapping
 * cv::Mat <-> raw pointers and mtfb_status -> mtf::utils::Exception.  Three classes:
 *
 *   mtf::b200::Tracker       : mtf::TrackerBase   one patch; what mtf::getTracker returns for sm "b200_fclk" ...
 *   mtf::b200::PFTracker     : mtf::TrackerBase   one object tracked by the particle filter (nt::PF, SM/src/NT/PF.cc)
 *   mtf::b200::Batch                              P patches tracked by ONE launch per frame
 *   mtf::b200::BatchMember   : mtf::TrackerBase   patch i of a Batch, so that composites that hold a
 *                                                 vector<TrackerBase*> (GridTracker, SM/src/GridTracker.cc:247-264)
 *                                                 drive the whole batch through their existing loop
 *
 * Interface replaced: include/mtf/TrackerBase.h:9-70 (setImage / initialize / update / setRegion / getRegion /
 * inputType), constructed where include/mtf/mtf.h:1282-1300 (getSM) constructs nt::FCLK / nt::ESM / nt::ICLK.
 * Ownership, threading and error behaviour follow SURVEY.md 8(b): raw `new`, caller owns; an instance is
 * single-threaded; errors are mtf::utils exceptions.
 */
#ifndef MTF_B200_TRACKER_H
#define MTF_B200_TRACKER_H

#include "mtf/TrackerBase.h"
#include "mtf/Utilities/excpUtils.h"
#include "opencv2/core/core.hpp"

#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "mtf_b200.h"

namespace mtf {
namespace b200 {

//! mtfb_status -> the exception the reference would have thrown (excpUtils.h:8-55)
inline void check(mtfb_status st){
	if(st == MTFB_OK){ return; }
	const std::string msg = std::string("mtf_b200 :: ") + mtfb_last_error();
	switch(st){
	case MTFB_ERR_INVALID_ARG: throw mtf::utils::InvalidArgument(msg);
	case MTFB_ERR_NOT_SUPPORTED: throw mtf::utils::FunctonNotImplemented(msg);
	case MTFB_ERR_INVALID_STATE: throw mtf::utils::InvalidTrackerState(msg);
	default: throw mtf::utils::LogicError(msg);
	}
}

//! parameters from the names MTF's factory uses ("fclk", "ssd", "8" ...; include/mtf/mtf.h:1066-1300)
inline mtfb_params makeParams(const char *sm, const char *am, const char *ssm, int n_patches, int resx, int resy){
	mtfb_params p;
	mtfb_default_params(&p);
	if(!strcmp(sm, "esm")){ p.sm = MTFB_SM_ESM; p.hess_type = MTFB_ESM_HESS_SUM_OF_SELF; }
	else if(!strcmp(sm, "fclk") || !strcmp(sm, "fc")){ p.sm = MTFB_SM_FCLK; p.hess_type = MTFB_LK_HESS_CURRENT_SELF; }
	else if(!strcmp(sm, "iclk") || !strcmp(sm, "ic")){ p.sm = MTFB_SM_ICLK; p.hess_type = MTFB_LK_HESS_INITIAL_SELF; }
	else if(!strcmp(sm, "falk") || !strcmp(sm, "fa")){ p.sm = MTFB_SM_FALK; p.hess_type = MTFB_LK_HESS_INITIAL_SELF; }
	else if(!strcmp(sm, "ialk") || !strcmp(sm, "ia")){ p.sm = MTFB_SM_IALK; p.hess_type = MTFB_LK_HESS_INITIAL_SELF; }
	else if(!strcmp(sm, "pf")){ p.sm = MTFB_SM_PF; }
	else{ throw mtf::utils::InvalidArgument(std::string("mtf_b200 :: unknown search method ") + sm); }
	if(!strcmp(am, "ssd")){ p.am = MTFB_AM_SSD; }
	else if(!strcmp(am, "ncc")){ p.am = MTFB_AM_NCC; }
	else if(!strcmp(am, "mi")){ p.am = MTFB_AM_MI; }
	else{ throw mtf::utils::InvalidArgument(std::string("mtf_b200 :: unknown appearance model ") + am); }
	if(!strcmp(ssm, "8") || !strcmp(ssm, "hom") || !strcmp(ssm, "homography")){ p.ssm = MTFB_SSM_HOMOGRAPHY; }
	else if(!strcmp(ssm, "6") || !strcmp(ssm, "aff") || !strcmp(ssm, "affine")){ p.ssm = MTFB_SSM_AFFINE; }
	else if(!strcmp(ssm, "2") || !strcmp(ssm, "trans") || !strcmp(ssm, "translation")){ p.ssm = MTFB_SSM_TRANSLATION; }
	else{ throw mtf::utils::InvalidArgument(std::string("mtf_b200 :: unknown state space model ") + ssm); }
	p.n_patches = n_patches; p.resx = resx; p.resy = resy;
	return p;
}

//! P patches sharing one image: thin RAII wrapper of mtfb_ctx
class Batch{
public:
	explicit Batch(const mtfb_params &params) : ctx(nullptr), prm(params), corners(8 * (size_t)params.n_patches, 0.0),
		n_supplied(0), frame_id(0), updated_frame(-1), gen(0), raw_channels(0), gauss_kernel_size(5), gauss_sigma(3.0){
		check(mtfb_create(&prm, &ctx));
	}
	//! make a batch created with sm "pf" a particle filter tracker (mtfb_pf_configure: nt::PF's constructor + initializeSampler)
	void configurePF(const mtfb_pf_params &pf){ check(mtfb_pf_configure(ctx, &pf)); }
	//! Take RAW uint8 frames (channels = 1 gray / 3 BGR) and run MTF's default pre-processing -- utils::GaussianSmoothing,
	//! Utilities/src/preprocUtils.cc:108-127, parameters.h:229-235 -- on the device behind the upload.  The application then
	//! creates this tracker's pre-processor with pre_proc_type "none" (inputType() reports CV_8UC1 / CV_8UC3).
	void useRawInput(int channels, int kernel_size = 5, double sigma = 3.0){
		if(channels != 1 && channels != 3){ throw mtf::utils::InvalidArgument("mtf_b200 :: useRawInput :: channels must be 1 or 3"); }
		raw_channels = channels; gauss_kernel_size = kernel_size; gauss_sigma = sigma;
	}
	int inputType() const{ return raw_channels == 3 ? CV_8UC3 : raw_channels == 1 ? CV_8UC1 : CV_32FC1; }
	~Batch(){ mtfb_destroy(ctx); }
	Batch(const Batch&) = delete;
	Batch& operator=(const Batch&) = delete;

	int size() const{ return prm.n_patches; }
	//! TrackerBase::setImage: the reference keeps the cv::Mat header and re-reads the pixels on every call
	//! (TrackerBase.h:21-26), so only the header is kept here and the upload happens in initialize()/update()
	void setImage(const cv::Mat &img){
		if(img.type() != inputType()){ throw mtf::utils::InvalidArgument("mtf_b200 :: setImage :: image type differs from inputType()"); }
		curr_img = img;
	}
	void upload(){
		if(curr_img.empty()){ throw mtf::utils::LogicError("mtf_b200 :: setImage has not been called"); }
		if(raw_channels){
			check(mtfb_set_image_u8(ctx, curr_img.ptr<unsigned char>(), curr_img.rows, curr_img.cols,
				static_cast<int>(curr_img.step), raw_channels, gauss_kernel_size, gauss_sigma));
		} else{
			check(mtfb_set_image(ctx, curr_img.ptr<float>(), curr_img.rows, curr_img.cols,
				static_cast<int>(curr_img.step / sizeof(float))));
		}
		++frame_id;
	}
	//! all P regions at once: corners = P x (2 x 4) doubles
	void initialize(const double *all_corners){ upload(); check(mtfb_initialize(ctx, all_corners)); fetch(); }
	void setRegion(const double *all_corners){ upload(); check(mtfb_set_region(ctx, all_corners)); fetch(); }
	void update(){
		if(n_supplied != 0){
			throw mtf::utils::LogicError("mtf_b200 :: update() while only some members of the batch have supplied their corners");
		}
		upload(); check(mtfb_update(ctx)); fetch(); updated_frame = frame_id;
	}
	// ---- GridTracker::update's step after its cells (SM/src/GridTracker.cc:265-280) on the device.
	//! `ssm.estimateWarpFromPts(ssm_update, pix_mask, prev_pts, curr_pts, est_params)` (GridTracker.cc:269) for the grid's own
	//! Homography / Affine: in_pts / out_pts = reinterpret_cast<const float*>(prev_pts.data()) of the std::vector<cv::Point2f>,
	//! state_update = ssm_update.data(), mask = pix_mask.data().  Returns the estimator's result (false: zero update).
	bool estimateWarpFromPts(int ssm, double *state_update, unsigned char *mask, const float *in_pts, const float *out_pts, int n,
		const mtfb_est_params &ep){
		int info[4];
		check(mtfb_estimate_warp_from_pts(ctx, ssm, in_pts, out_pts, n, &ep, state_update, mask, nullptr, info));
		return info[0] != 0;
	}
	//! the same without the points leaving the device: prev_pts / curr_pts are the centroids of this batch's regions at the last
	//! initialize / setRegion / gridCommit and now (GridTracker.cc:257, :389); call gridEnable() once after construction
	void gridEnable(){ check(mtfb_grid_enable(ctx)); }
	bool gridEstimate(int ssm, double *state_update, unsigned char *mask, const mtfb_est_params &ep){
		int info[4];
		check(mtfb_grid_estimate(ctx, ssm, &ep, state_update, mask, nullptr, info));
		return info[0] != 0;
	}
	//! prev_pts = curr_pts (reset_at_each_frame = 0, GridTracker.cc:276-279)
	void gridCommit(){ check(mtfb_grid_commit(ctx)); }
	//! everything GridTracker::update does after its cells, queued on the device without a host hop (GridTracker.cc:265-280 +
	//! resetTrackers(true), :345-392, for the shipped grid: reset_at_each_frame = 1, patch_centroid_inside = 1, fixed patch size):
	//! the estimate, `ssm.applyWarpToCorners` + `ssm.setCorners` on `region` (2 x 4 doubles of the grid's own SSM, updated in
	//! place), every cell re-initialised at the patch_size box around the centroid of its four grid points.  A GridBase subclass
	//! that owns this Batch calls it from update() and copies `region` into cv_corners_mat.
	//! GridTracker::initialize / setRegion of the shipped grid: the cells' regions for `region`, every cell initialised there
	void gridInitialize(int grid_size_x, int grid_size_y, double patch_size_x, double patch_size_y, const double *region){
		upload();
		check(mtfb_grid_initialize(ctx, grid_size_x, grid_size_y, patch_size_x, patch_size_y, region));
		fetch();
	}
	bool gridAdvance(int ssm, int grid_size_x, int grid_size_y, double patch_size_x, double patch_size_y, double *region,
		double *state_update, unsigned char *mask, const mtfb_est_params &ep){
		int info[4];
		check(mtfb_grid_advance(ctx, ssm, &ep, grid_size_x, grid_size_y, patch_size_x, patch_size_y, region, state_update, mask, nullptr, info));
		fetch();
		return info[0] != 0;
	}
	//! mtf::SSMEstimatorParams -> mtfb_est_params (method: SSMEstimatorParams::EstType as an int); seed replaces random_device
	static mtfb_est_params estParams(int method, double ransac_reproj_thresh, int n_model_pts, bool refine, int max_iters,
		int max_subset_attempts, double confidence, int lm_max_iters, unsigned long long seed){
		mtfb_est_params ep;
		mtfb_est_default_params(&ep);
		ep.method = method; ep.ransac_reproj_thresh = ransac_reproj_thresh; ep.n_model_pts = n_model_pts; ep.refine = refine ? 1 : 0;
		ep.max_iters = max_iters; ep.max_subset_attempts = max_subset_attempts; ep.confidence = confidence;
		ep.lm_max_iters = lm_max_iters; ep.seed = seed;
		return ep;
	}
	// ---- the batch split over several GPUs (one process each): the per-frame exchange of the regions over NVLink peer memory.
	//! peerExport: this rank's gathered array for a job of n_total patches -> its 64-byte handle, to be exchanged by the host
	//! (MPI_Allgather, a pipe, a file); peerAttach: rank / world / this rank's first patch / all the handles; from then on
	//! update() stores every region into the arrays of all ranks and peerGather() (once per frame, on every rank) completes
	//! the exchange: gathered[i] = the region of patch i of the whole job
	mtfb_peer_handle peerExport(int n_total){
		mtfb_peer_handle h;
		check(mtfb_peer_export(ctx, n_total, &h));
		gathered.assign(8 * (size_t)n_total, 0.0);
		return h;
	}
	void peerAttach(int rank, int world, int row0, const mtfb_peer_handle *handles){ check(mtfb_peer_attach(ctx, rank, world, row0, handles)); }
	const double* peerGather(){
		check(mtfb_peer_gather(ctx));
		check(mtfb_get_gathered_region(ctx, gathered.data()));
		return gathered.data();
	}
	//! bumped whenever the regions are refreshed from the device (initialize / setRegion / update)
	long generation() const{ return gen; }
	const double* region(int i) const{ return &corners[8 * (size_t)i]; }
	mtfb_ctx* handle(){ return ctx; }

	// ---- used by BatchMember: members hand in their corners one by one; the batch is (re)initialised when the
	// last one has arrived
	void supplyCorners(int i, const cv::Mat &c, bool is_init){
		for(int k = 0; k < 8; ++k){ pending_corners()[8 * (size_t)i + k] = c.at<double>(k / 4, k % 4); }
		if(++n_supplied == size()){
			n_supplied = 0;
			if(is_init){ initialize(pending.data()); } else{ setRegion(pending.data()); }
		}
	}
	const cv::Mat& image() const{ return curr_img; }

private:
	std::vector<double>& pending_corners(){ if(pending.empty()){ pending.assign(corners.size(), 0.0); } return pending; }
	void fetch(){ check(mtfb_get_corners(ctx, corners.data())); ++gen; }
	mtfb_ctx *ctx;
	mtfb_params prm;
	cv::Mat curr_img;
	std::vector<double> corners, pending, gathered;
	int n_supplied, frame_id, updated_frame;
	long gen;
	int raw_channels, gauss_kernel_size;
	double gauss_sigma;
};

//! One patch tracked on the GPU; drop-in for nt::FCLK / nt::ESM / nt::ICLK (SM/src/NT/*.cc)
class Tracker : public mtf::TrackerBase{
public:
	Tracker(const char *sm, const char *am, const char *ssm, int resx, int resy) :
		batch(makeParams(notPF(sm), am, ssm, 1, resx, resy)){
		name = std::string("b200_") + sm;
		cv_corners_mat.create(2, 4, CV_64FC1);
	}
	explicit Tracker(const mtfb_params &params) : batch(params){
		name = "b200";
		cv_corners_mat.create(2, 4, CV_64FC1);
	}
	using TrackerBase::initialize;
	using TrackerBase::update;
	using TrackerBase::setRegion;
	void setImage(const cv::Mat &img) override{ batch.setImage(img); }
	void initialize(const cv::Mat &corners) override{ toArray(corners); batch.initialize(c8); publish(); }
	void update() override{ batch.update(); publish(); }
	void setRegion(const cv::Mat &corners) override{ toArray(corners); batch.setRegion(c8); publish(); }
	int inputType() const override{ return batch.inputType(); }
	Batch& getBatch(){ return batch; }
protected:
	//! the particle filter needs its sampler (PFParams): PFTracker below
	static const char* notPF(const char *sm){
		if(!strcmp(sm, "pf")){ throw mtf::utils::InvalidArgument("mtf_b200 :: sm \"pf\" is constructed as mtf::b200::PFTracker (it needs PFParams)"); }
		return sm;
	}
private:
	void toArray(const cv::Mat &corners){
		if(corners.rows != 2 || corners.cols != 4 || corners.type() != CV_64FC1){
			throw mtf::utils::InvalidArgument("mtf_b200 :: corners must be a 2 x 4 CV_64FC1 matrix");
		}
		for(int k = 0; k < 8; ++k){ c8[k] = corners.at<double>(k / 4, k % 4); }
	}
	void publish(){
		const double *r = batch.region(0);
		for(int k = 0; k < 8; ++k){ cv_corners_mat.at<double>(k / 4, k % 4) = r[k]; }
	}
	Batch batch;
	double c8[8];
};

//! One object tracked by the particle filter on the GPU; drop-in for nt::PF (SM/src/NT/PF.cc), constructed where
//! include/mtf/mtf.h:345-346 constructs PF<AM, SSM>(getPFParams().get(), ...): the caller fills mtfb_pf_params from PFParams
//! (n_particles, dynamic_model, update_type, likelihood_func, resampling_type, mean_type, reset_to_mean,
//! adaptive_resampling_thresh, measurement_sigma, the first distribution's ssm_sigma / ssm_mean) and
//! HomographyParams::corner_based_sampling.
class PFTracker : public mtf::TrackerBase{
public:
	PFTracker(const char *am, const char *ssm, int resx, int resy, const mtfb_pf_params &pf) :
		batch(makeParams("pf", am, ssm, 1, resx, resy)){
		batch.configurePF(pf);
		name = "b200_pf";
		cv_corners_mat.create(2, 4, CV_64FC1);
	}
	using TrackerBase::initialize;
	using TrackerBase::update;
	using TrackerBase::setRegion;
	void setImage(const cv::Mat &img) override{ batch.setImage(img); }
	void initialize(const cv::Mat &corners) override{ toArray(corners); batch.initialize(c8); publish(); }
	void update() override{ batch.update(); publish(); }
	void setRegion(const cv::Mat &corners) override{ toArray(corners); batch.setRegion(c8); publish(); }
	int inputType() const override{ return batch.inputType(); }
	Batch& getBatch(){ return batch; }
private:
	void toArray(const cv::Mat &corners){
		if(corners.rows != 2 || corners.cols != 4 || corners.type() != CV_64FC1){
			throw mtf::utils::InvalidArgument("mtf_b200 :: corners must be a 2 x 4 CV_64FC1 matrix");
		}
		for(int k = 0; k < 8; ++k){ c8[k] = corners.at<double>(k / 4, k % 4); }
	}
	void publish(){
		const double *r = batch.region(0);
		for(int k = 0; k < 8; ++k){ cv_corners_mat.at<double>(k / 4, k % 4) = r[k]; }
	}
	Batch batch;
	double c8[8];
};

//! mtf::GridTracker<SSM> (SM/include/mtf/SM/GridTracker.h, SM/src/GridTracker.cc) for the shipped grid -- grid_size_x x grid_size_y
//! patch trackers of ONE search method / appearance model / cell SSM, reset_at_each_frame = 1, patch_centroid_inside = 1, fixed
//! patch size, no forward-backward estimation -- as ONE batch on the device: the cells' update is one launch, and everything
//! after it (the robust estimate of the grid's own Homography / Affine from the cells' centroids, the region's new corners, the
//! cells' next regions, their re-initialisation) is queued behind it without a host hop (Batch::gridAdvance).  The reference's
//! GridTracker is composed in mtf.h:748-802 from a vector<TrackerBase*>; this class takes its place there when every cell runs
//! the same tracker.  est: Batch::estParams(...); its seed advances by one per frame (the reference seeds every frame's
//! estimator from random_device, SSMEstimator.cc:22-24).
class GridTracker : public mtf::TrackerBase{
public:
	GridTracker(const char *sm, const char *am, const char *cell_ssm, int grid_size_x, int grid_size_y, int patch_size_x, int patch_size_y,
		int cell_resx, int cell_resy, const char *grid_ssm, const mtfb_est_params &est) :
		batch(makeParams(sm, am, cell_ssm, grid_size_x*grid_size_y, cell_resx, cell_resy)), gx(grid_size_x), gy(grid_size_y),
		psx(patch_size_x), psy(patch_size_y), ep(est), seed0(est.seed), frame(0), mask((size_t)grid_size_x*grid_size_y, 1){
		if(!strcmp(grid_ssm, "8") || !strcmp(grid_ssm, "hom") || !strcmp(grid_ssm, "homography")){ ssm = MTFB_SSM_HOMOGRAPHY; }
		else if(!strcmp(grid_ssm, "6") || !strcmp(grid_ssm, "aff") || !strcmp(grid_ssm, "affine")){ ssm = MTFB_SSM_AFFINE; }
		else{ throw mtf::utils::InvalidArgument(std::string("mtf_b200 :: GridTracker: the grid's SSM must be homography or affine, not ") + grid_ssm); }
		name = "b200_grid";
		cv_corners_mat.create(2, 4, CV_64FC1);
		batch.gridEnable();
	}
	using TrackerBase::initialize;
	using TrackerBase::update;
	using TrackerBase::setRegion;
	void setImage(const cv::Mat &img) override{ batch.setImage(img); }
	//! GridTracker::initialize (GridTracker.cc:232-245): ssm.initialize(corners); resetTrackers(true)
	void initialize(const cv::Mat &corners) override{ toRegion(corners); batch.gridInitialize(gx, gy, psx, psy, region); publish(); }
	//! GridTracker::update (GridTracker.cc:247-285)
	void update() override{
		batch.update();
		ep.seed = seed0 + (unsigned long long)(++frame);
		batch.gridAdvance(ssm, gx, gy, psx, psy, region, ssm_update, mask.data(), ep);
		publish();
	}
	//! GridTracker::setRegion (GridTracker.cc:287-291): ssm.setCorners(corners); resetTrackers(reinit_at_each_frame = true)
	void setRegion(const cv::Mat &corners) override{ toRegion(corners); batch.gridInitialize(gx, gy, psx, psy, region); publish(); }
	int inputType() const override{ return batch.inputType(); }
	//! the estimator's inlier mask of the last update (pix_mask, GridTracker.h) and the cells' regions
	const std::vector<unsigned char>& getPixMask() const{ return mask; }
	Batch& getBatch(){ return batch; }
private:
	void toRegion(const cv::Mat &corners){
		if(corners.rows != 2 || corners.cols != 4 || corners.type() != CV_64FC1){
			throw mtf::utils::InvalidArgument("mtf_b200 :: corners must be a 2 x 4 CV_64FC1 matrix");
		}
		for(int k = 0; k < 8; ++k){ region[k] = corners.at<double>(k / 4, k % 4); }
	}
	void publish(){ for(int k = 0; k < 8; ++k){ cv_corners_mat.at<double>(k / 4, k % 4) = region[k]; } }
	Batch batch;
	int gx, gy, ssm;
	double psx, psy;
	mtfb_est_params ep;
	unsigned long long seed0;
	long frame;
	double region[8], ssm_update[8];
	std::vector<unsigned char> mask;
};

//! Patch i of a shared Batch.  A composite that loops `trackers[i]->update()` (GridTracker.cc:256-259, serial
//! build) triggers one launch for the whole batch on member 0 and reads results for the others.
class BatchMember : public mtf::TrackerBase{
public:
	BatchMember(std::shared_ptr<Batch> _batch, int _id) : batch(_batch), id(_id), pending(false), supplied_gen(-1){
		name = "b200_member";
		cv_corners_mat.create(2, 4, CV_64FC1);
	}
	using TrackerBase::initialize;
	using TrackerBase::update;
	using TrackerBase::setRegion;
	void setImage(const cv::Mat &img) override{ if(id == 0){ batch->setImage(img); } }
	//! The batch is (re)initialised when its LAST member has supplied corners; until then this member's region is the one it
	//! supplied -- GridTracker::resetTrackers reads trackers[i]->getRegion() right after trackers[i]->initialize(corners)
	//! (SM/src/GridTracker.cc:381-387), and after setCorners the region equals the supplied corners exactly
	void initialize(const cv::Mat &corners) override{ supply(corners, true); }
	void setRegion(const cv::Mat &corners) override{ supply(corners, false); }
	//! member 0 launches the whole batch; the composite's loop visits the members in index order
	//! (GridTracker.cc:256-259; the TBB / OpenMP variants of that loop, :248-255, must stay disabled)
	void update() override{ if(id == 0){ batch->update(); } }
	const cv::Mat& getRegion() override{
		if(pending && batch->generation() == supplied_gen){ return cv_corners_mat; }      // not flushed to the device yet
		pending = false;
		const double *r = batch->region(id);
		for(int k = 0; k < 8; ++k){ cv_corners_mat.at<double>(k / 4, k % 4) = r[k]; }
		return cv_corners_mat;
	}
	int inputType() const override{ return batch->inputType(); }
private:
	void supply(const cv::Mat &corners, bool is_init){
		if(corners.rows != 2 || corners.cols != 4 || corners.type() != CV_64FC1){
			throw mtf::utils::InvalidArgument("mtf_b200 :: corners must be a 2 x 4 CV_64FC1 matrix");
		}
		corners.copyTo(cv_corners_mat);
		pending = true; supplied_gen = batch->generation();
		batch->supplyCorners(id, corners, is_init);
	}
	std::shared_ptr<Batch> batch;
	int id;
	bool pending;
	long supplied_gen;
};

//! grid_res^2 members over one batch: what replaces the loop at include/mtf/mtf.h:786-789
inline std::vector<mtf::TrackerBase*> makeBatchMembers(const char *sm, const char *am, const char *ssm,
	int n_trackers, int resx, int resy){
	std::shared_ptr<Batch> batch(new Batch(makeParams(sm, am, ssm, n_trackers, resx, resy)));
	std::vector<mtf::TrackerBase*> out;
	for(int i = 0; i < n_trackers; ++i){ out.push_back(new BatchMember(batch, i)); }
	return out;
}

} // namespace b200
} // namespace mtf

#endif
