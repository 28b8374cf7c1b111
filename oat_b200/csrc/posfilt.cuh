// posfilt.cuh -- the O(1) epilogue behind the detectors: per-source 2-D
// Kalman filter (oat posifilt kalman: src/positionfilter/KalmanFilter2D.cpp:95-200 on
// cv::KalmanFilter(4, 2, 0, CV_64F)) and the mean combiner (oat posicom mean:
// src/positioncombiner/MeanPosition.cpp:60-118), one launch of one warp: lane i filters source i,
// lane 0 combines.  fp64 like the reference.  The state is [x x' y y'], the transition matrix is
// identity plus dt at (0,1) and (2,3), and the measurement matrix picks rows 0 and 2, so the matrix
// products of cv::KalmanFilter::predict/correct are written out for that structure (the covariance
// stays a full 4x4: nothing assumes x and y decouple).  The reference's observable behaviour is kept,
// including that it reports the PREDICTED state, that re-initialisation leaves errorCovPost alone, and
// that timeout 0 (the default) never validates a position (KalmanFilter2D.cpp:101-117).
#pragma once
#include "common.cuh"

namespace oat {

struct KalmanConsts {
    double dt, q00, q01, q11, r;  // process-noise block [[q00 q01][q01 q11]] (both axes), measurement variance
    int not_found_thr;            // int(timeout / dt)
    int enabled;
};

struct KalmanState {
    double P[4][4];      // errorCovPost
    double post[4];      // statePost
    double predicted[4]; // last predict() result (what the reference publishes)
    double meas[2];      // last valid measurement
    int found, not_found;
};

struct PosfiltDev {       // device-resident part of an oat_posfilt
    KalmanState k[8];
    int stalled;          // a source frame's detection was not final (tail overflow): updates are held
};

__device__ inline void kalman_reset(KalmanState &k)
{
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 4; ++j) k.P[i][j] = 0.0;
        k.post[i] = 0.0;
        k.predicted[i] = 6.0;  // cv::Mat_<double>{4, 1, CV_64F}: KalmanFilter2D.h:63 (flagged invalid anyway)
    }
    k.meas[0] = k.meas[1] = 6.0;
    k.found = 0;
    k.not_found = 0;
}

// x = pinv(S) for a symmetric 2x2 (cv::solve(..., DECOMP_SVD) semantics: singular values below
// 2*eps*sum count as zero), through the eigen-decomposition
__device__ inline void pinv2(double a, double b, double d, double &ia, double &ib, double &id)
{
    const double tr = a + d, df = a - d;
    const double rad = sqrt(df * df + 4.0 * b * b);
    const double l1 = 0.5 * (tr + rad), l2 = 0.5 * (tr - rad);
    double ux = 1.0, uy = 0.0;
    if (b != 0.0) {
        ux = l1 - d;
        uy = b;
        const double n = sqrt(ux * ux + uy * uy);
        ux /= n;
        uy /= n;
    } else if (a < d) {
        ux = 0.0;
        uy = 1.0;
    }
    const double thr = 2.0 * 2.220446049250313e-16 * (fabs(l1) + fabs(l2));
    const double i1 = fabs(l1) > thr ? 1.0 / l1 : 0.0, i2 = fabs(l2) > thr ? 1.0 / l2 : 0.0;
    ia = i1 * ux * ux + i2 * uy * uy;
    ib = (i1 - i2) * ux * uy;
    id = i1 * uy * uy + i2 * ux * ux;
}

// KalmanFilter2D::filter: p carries the raw measurement in, the filtered position/velocity out
__device__ inline void kalman_filter(KalmanState &k, const KalmanConsts &c, oat_position &p)
{
    if (p.position_valid) {
        k.meas[0] = p.x;
        k.meas[1] = p.y;
        k.not_found = 0;
        if (!k.found) {  // initializeFilter: state from the measurement; errorCovPost is NOT touched
            k.post[0] = p.x;
            k.post[1] = 0.0;
            k.post[2] = p.y;
            k.post[3] = 0.0;
        }
        k.found = 1;
    } else {
        ++k.not_found;
    }
    if (k.not_found >= c.not_found_thr) k.found = 0;
    if (k.found) {
        const double dt = c.dt;
        // predict: pre = A post;  Ppre = A P A' + Q
        double pre[4] = {k.post[0] + dt * k.post[1], k.post[1], k.post[2] + dt * k.post[3], k.post[3]};
        double T[4][4], M[4][4];
        for (int j = 0; j < 4; ++j) {
            T[0][j] = k.P[0][j] + dt * k.P[1][j];
            T[1][j] = k.P[1][j];
            T[2][j] = k.P[2][j] + dt * k.P[3][j];
            T[3][j] = k.P[3][j];
        }
        for (int i = 0; i < 4; ++i) {
            M[i][0] = T[i][0] + T[i][1] * dt;
            M[i][1] = T[i][1];
            M[i][2] = T[i][2] + T[i][3] * dt;
            M[i][3] = T[i][3];
        }
        M[0][0] += c.q00;
        M[0][1] += c.q01;
        M[1][0] += c.q01;
        M[1][1] += c.q11;
        M[2][2] += c.q00;
        M[2][3] += c.q01;
        M[3][2] += c.q01;
        M[3][3] += c.q11;
        for (int i = 0; i < 4; ++i) k.predicted[i] = pre[i];
        // correct: H picks rows 0 and 2.  S = H Ppre H' + R;  G' = pinv(S) (H Ppre)
        const double s00 = M[0][0] + c.r, s01 = 0.5 * (M[0][2] + M[2][0]), s11 = M[2][2] + c.r;
        double ia, ib, id;
        pinv2(s00, s01, s11, ia, ib, id);
        double G0[4], G1[4];
        for (int i = 0; i < 4; ++i) {
            G0[i] = ia * M[0][i] + ib * M[2][i];
            G1[i] = ib * M[0][i] + id * M[2][i];
        }
        const double e0 = k.meas[0] - pre[0], e1 = k.meas[1] - pre[2];
        for (int i = 0; i < 4; ++i) k.post[i] = pre[i] + (G0[i] * e0 + G1[i] * e1);
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) k.P[i][j] = M[i][j] - (G0[i] * M[0][j] + G1[i] * M[2][j]);
    }
    p.x = k.predicted[0];
    p.vx = k.predicted[1];
    p.y = k.predicted[2];
    p.vy = k.predicted[3];
    p.position_valid = p.velocity_valid = k.found ? 1 : 0;
}

struct PosfiltArgs {
    PosfiltDev *dev;
    KalmanConsts kc;
    int n;                       // sources (<= 8)
    int combine;                 // 1: MeanPosition::combine over the n (filtered) sources; 0: out[i] = source i
    int heading_anchor;          // < 0: headings are averaged, not generated
    const oat_position *raw;     // n raw positions, or NULL when `det` feeds source 0
    const oat_detection *det;    // a detector's device-resident result (the tracker's epilogue)
    const int32_t *det_status;   // or NULL: != 0 means *det is not final yet -> hold the filter (stall)
    int force;                   // clear a stall first (the host has made *det final)
    oat_position *out;           // combine ? 1 : n positions
};

__global__ void posfilt_kernel(const PosfiltArgs a)
{
    __shared__ oat_position sp[8];
    const int i = threadIdx.x;
    if (a.force && i == 0) a.dev->stalled = 0;
    __syncwarp();
    if (!a.force && ((a.det_status && *a.det_status != 0) || a.dev->stalled)) {
        if (i == 0) {
            a.dev->stalled = 1;
            a.out[0].reserved = -1;  // tells the host to re-run this frame once its detection is final
        }
        return;
    }
    if (i < a.n) {
        oat_position p;
        if (a.raw) {
            p = a.raw[i];
        } else {
            p = oat_position{};
            p.position_valid = a.det->position_valid;
            p.x = a.det->x;
            p.y = a.det->y;
        }
        p.reserved = 0;
        if (a.kc.enabled) kalman_filter(a.dev->k[i], a.kc, p);
        sp[i] = p;
        if (!a.combine) a.out[i] = p;
    }
    __syncwarp();
    if (a.combine && i == 0) {
        const double md = 1.0 / (double)a.n;
        oat_position o{};
        o.position_valid = o.velocity_valid = o.heading_valid = 1;
        for (int s = 0; s < a.n; ++s) {
            const oat_position &p = sp[s];
            if (p.position_valid) {
                o.x += md * p.x;
                o.y += md * p.y;
            } else
                o.position_valid = 0;
            if (p.velocity_valid) {
                o.vx += md * p.vx;
                o.vy += md * p.vy;
            } else
                o.velocity_valid = 0;
            if (a.heading_anchor >= 0) {  // anchor -> source vectors, while every position so far was valid
                if (o.position_valid) {
                    o.hx += p.x - sp[a.heading_anchor].x;
                    o.hy += p.y - sp[a.heading_anchor].y;
                } else
                    o.heading_valid = 0;
            } else if (p.heading_valid) {
                o.hx += p.hx;
                o.hy += p.hy;
            } else
                o.heading_valid = 0;
        }
        if (o.heading_valid) {
            const double mag = sqrt(o.hx * o.hx + o.hy * o.hy);
            o.hx = o.hx / mag;
            o.hy = o.hy / mag;
        }
        a.out[0] = o;
    }
}

__global__ void posfilt_reset_kernel(PosfiltDev *d)
{
    if (threadIdx.x < 8) kalman_reset(d->k[threadIdx.x]);
    if (threadIdx.x == 0) d->stalled = 0;
}

}  // namespace oat
