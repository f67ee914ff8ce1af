// eigen_ref.cpp -- TEST INFRASTRUCTURE ONLY.
// Thin extern "C" harness around the REAL Eigen 3.2.92 vendored by the reference
// (/root/reference/lidar_localization/third_party/eigen3), compiled from the headers where they lie
// by oracle/Makefile into oracle/_ref/libeigen_ref.so.  It evaluates exactly the Eigen expressions
// that pcl::NormalDistributionsTransform / pcl::VoxelGridCovariance 1.7 (and the in-tree copy,
// NDTM/NormalDistributionsTransform.cpp:331-337,353-355,370-373; NDTM/VoxelGrid.cpp:293-320) use,
// so the plain-C restatements in ndt_oracle.c can be pinned against them.
#include <Eigen/Dense>
#include <Eigen/Geometry>
#include <limits>

extern "C" {

int ref_svd_solve6(const double *H, const double *b, double *x, double *sv) {
    Eigen::Matrix<double, 6, 6> hessian = Eigen::Map<const Eigen::Matrix<double, 6, 6> >(H);
    Eigen::Matrix<double, 6, 1> g = Eigen::Map<const Eigen::Matrix<double, 6, 1> >(b);
    Eigen::JacobiSVD<Eigen::Matrix<double, 6, 6> > svd(hessian, Eigen::ComputeFullU | Eigen::ComputeFullV);
    Eigen::Matrix<double, 6, 1> r = svd.solve(g);
    for (int i = 0; i < 6; ++i) { x[i] = r(i); sv[i] = svd.singularValues()(i); }
    return (int)svd.rank();
}

void ref_euler012(const float *T, float *out) {
    Eigen::Matrix4f m = Eigen::Map<const Eigen::Matrix4f>(T);
    Eigen::Transform<float, 3, Eigen::Affine, Eigen::ColMajor> eig;
    eig.matrix() = m;
    Eigen::Vector3f r = eig.rotation().eulerAngles(0, 1, 2);
    out[0] = r(0); out[1] = r(1); out[2] = r(2);
}

void ref_pose_matrix(const double *p, float *T) {
    Eigen::Matrix4f m = (Eigen::Translation<float, 3>(static_cast<float>(p[0]), static_cast<float>(p[1]), static_cast<float>(p[2])) *
                         Eigen::AngleAxis<float>(static_cast<float>(p[3]), Eigen::Vector3f::UnitX()) *
                         Eigen::AngleAxis<float>(static_cast<float>(p[4]), Eigen::Vector3f::UnitY()) *
                         Eigen::AngleAxis<float>(static_cast<float>(p[5]), Eigen::Vector3f::UnitZ())).matrix();
    for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) T[c * 4 + r] = m(r, c);
}

// second pass of VoxelGridCovariance::applyFilter for one leaf with n >= min_points:
// inputs are the raw accumulators (pt_sum, cov_ accumulated from Identity), row-major 3x3.
// returns nr_points (n or -1)
int ref_leaf_finish(const double *pt_sum_in, const double *cov_acc_in, int n, double eig_mult,
                    double *mean_out, double *cov_out, double *icov_out, double *evals_out) {
    Eigen::Vector3d pt_sum(pt_sum_in[0], pt_sum_in[1], pt_sum_in[2]);
    Eigen::Matrix3d cov_;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) cov_(r, c) = cov_acc_in[r * 3 + c];
    Eigen::Vector3d mean_ = pt_sum;
    int nr_points = n;
    mean_ /= nr_points;
    Eigen::Matrix3d icov_ = Eigen::Matrix3d::Zero();
    Eigen::SelfAdjointEigenSolver<Eigen::Matrix3d> eigensolver;
    Eigen::Matrix3d eigen_val;
    Eigen::Matrix3d evecs_;
    cov_ = (cov_ - 2 * (pt_sum * mean_.transpose())) / nr_points + mean_ * mean_.transpose();
    cov_ *= (nr_points - 1.0) / nr_points;
    eigensolver.compute(cov_);
    eigen_val = eigensolver.eigenvalues().asDiagonal();
    evecs_ = eigensolver.eigenvectors();
    int ret = nr_points;
    if (eigen_val(0, 0) < 0 || eigen_val(1, 1) < 0 || eigen_val(2, 2) <= 0) {
        ret = -1;
    } else {
        double min_covar_eigvalue = eig_mult * eigen_val(2, 2);
        if (eigen_val(0, 0) < min_covar_eigvalue) {
            eigen_val(0, 0) = min_covar_eigvalue;
            if (eigen_val(1, 1) < min_covar_eigvalue) eigen_val(1, 1) = min_covar_eigvalue;
            cov_ = evecs_ * eigen_val * evecs_.inverse();
        }
        icov_ = cov_.inverse();
        if (icov_.maxCoeff() == std::numeric_limits<float>::infinity() ||
            icov_.minCoeff() == -std::numeric_limits<float>::infinity())
            ret = -1;
    }
    for (int r = 0; r < 3; ++r) {
        mean_out[r] = mean_(r);
        evals_out[r] = eigen_val(r, r);
        for (int c = 0; c < 3; ++c) { cov_out[r * 3 + c] = cov_(r, c); icov_out[r * 3 + c] = icov_(r, c); }
    }
    return ret;
}

void ref_inverse3(const double *A, double *out) {
    Eigen::Matrix3d m;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) m(r, c) = A[r * 3 + c];
    Eigen::Matrix3d inv = m.inverse();
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) out[r * 3 + c] = inv(r, c);
}

// pcl::transformPointCloud dense branch written with Eigen's Transform accessor
void ref_transform_point(const float *T, const float *p, float *out) {
    Eigen::Transform<float, 3, Eigen::Affine> transform(Eigen::Map<const Eigen::Matrix4f>(T).eval());
    Eigen::Matrix<float, 3, 1> pt(p[0], p[1], p[2]);
    out[0] = static_cast<float>(transform(0, 0) * pt.coeffRef(0) + transform(0, 1) * pt.coeffRef(1) + transform(0, 2) * pt.coeffRef(2) + transform(0, 3));
    out[1] = static_cast<float>(transform(1, 0) * pt.coeffRef(0) + transform(1, 1) * pt.coeffRef(1) + transform(1, 2) * pt.coeffRef(2) + transform(1, 3));
    out[2] = static_cast<float>(transform(2, 0) * pt.coeffRef(0) + transform(2, 1) * pt.coeffRef(1) + transform(2, 2) * pt.coeffRef(2) + transform(2, 3));
}
}
