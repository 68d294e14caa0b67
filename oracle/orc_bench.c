/*
 * orc_bench.c -- CPU-baseline harness around the oracle (TEST / BENCH INFRASTRUCTURE ONLY).
 *
 * Drives the restated reference loops of ouster_oracle.c the way the reference's own value-returning
 * entry points are called, from C, so that the timed region contains the reference's algorithm and
 * its per-call allocations and nothing of a Python harness.  Only bench.py's `cpu_baseline` leg and
 * `bench.py --impl reference` call this file.
 *
 * Three ways of running the same work, all in one process:
 *   mode 0  as shipped: one thread; destagger<uint32_t>() then cartesian() as separate calls, each
 *           returning a freshly allocated image / point matrix
 *           (impl/lidar_frame_impl.h:825-834 -> :733-760; xyzlut.h:139-150, impl/cartesian.h:81-90).
 *   mode 1  the reference's opt-in OpenMP build (-DOUSTER_OMP, impl/cartesian.h:15-23,50-52):
 *           `#pragma omp parallel for schedule(static)` over the pixels inside cartesianT, frames one
 *           after the other, everything else serial.
 *   mode 2  one thread per independent sensor stream (frames are independent; each thread runs the
 *           as-shipped single-thread code on its own frames) -- how a multi-sensor host uses the
 *           reference (sensor_frame_set_source.cpp:177-223: one FrameBatcher per sensor).
 *
 * All file:line citations are relative to /root/reference (ouster-sdk 1.0.1).
 */
#define _POSIX_C_SOURCE 200809L
#include <omp.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "ouster_oracle.h"

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int orc_bench_max_threads(void) { return omp_get_max_threads(); }

/* destagger<uint32_t>(img, shifts): fresh h x w image per call (impl/lidar_frame_impl.h:825-834) */
static uint32_t* destagger_new(const uint32_t* img, const int* shifts, size_t h, size_t w) {
    uint32_t* out = (uint32_t*)malloc(h * w * sizeof(uint32_t));
    if (out) orc_destagger(sizeof(uint32_t), 1, img, shifts, h, h, w, 0, out);
    return out;
}

/* cartesian(range, lut): fresh (h*w) x 3 matrix per call (impl/cartesian.h:81-90) */
static void* cartesian_new(const uint32_t* rng, const void* dir, const void* off, size_t n, int f64,
                           int omp) {
    void* pts = malloc(n * 3 * (f64 ? sizeof(double) : sizeof(float)));
    if (!pts) return NULL;
    if (f64) {
        if (omp) orc_cartesian_f64_omp((double*)pts, rng, (const double*)dir, (const double*)off, n);
        else orc_cartesian_f64((double*)pts, rng, (const double*)dir, (const double*)off, n);
    } else {
        if (omp) orc_cartesian_f32_omp((float*)pts, rng, (const float*)dir, (const float*)off, n);
        else orc_cartesian_f32((float*)pts, rng, (const float*)dir, (const float*)off, n);
    }
    return pts;
}

static double k1_frame(const uint32_t* rng, size_t R, size_t h, size_t w, const int* shifts,
                       const void* dir, const void* off, int f64, int omp) {
    double acc = 0.0;
    for (size_t r = 0; r < R; ++r) {
        const uint32_t* img = rng + r * h * w;
        uint32_t* rd = destagger_new(img, shifts, h, w);
        void* pts = cartesian_new(img, dir, off, h * w, f64, omp);
        if (rd && pts) /* touch the results so the calls cannot be elided */
            acc += (double)rd[h * w / 2] + (f64 ? ((double*)pts)[h * w] : (double)((float*)pts)[h * w]);
        free(rd);
        free(pts);
    }
    return acc;
}

/* K1 baseline: `reps` passes over F frames of R returns (rng = [F][R][h][w]); returns seconds.
 * threads <= 0: all cores.  *sink receives a value derived from the outputs. */
double orc_bench_k1(int mode, int f64, const uint32_t* rng, size_t F, size_t R, size_t h, size_t w,
                    const int* shifts, const void* dir, const void* off, int threads, int reps,
                    double* sink) {
    if (threads <= 0) threads = omp_get_max_threads();
    double acc = 0.0;
    const size_t fs = R * h * w;
    const double t0 = now_s();
    for (int rep = 0; rep < reps; ++rep) {
        if (mode == 2) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(+ : acc)
            for (ptrdiff_t f = 0; f < (ptrdiff_t)F; ++f)
                acc += k1_frame(rng + (size_t)f * fs, R, h, w, shifts, dir, off, f64, 0);
        } else {
            if (mode == 1) omp_set_num_threads(threads);
            for (size_t f = 0; f < F; ++f)
                acc += k1_frame(rng + f * fs, R, h, w, shifts, dir, off, f64, mode == 1);
        }
    }
    const double t = now_s() - t0;
    if (sink) *sink = acc;
    return t;
}

/* one frame of the packet path: FrameBatcher block/column parse of every packet into a LidarFrame
 * (lidar_frame.cpp:1698-1959), then destagger + cartesian of each range field */
static double k2_frame(const orc_packet_format* pf, orc_frame* fr, const uint8_t* packets, size_t n_pk,
                       size_t psz, const int* shifts, const void* dir, const void* off, int f64,
                       int omp) {
    double acc = 0.0;
    orc_batcher* b = orc_batcher_create(pf, 0, 0, pf->columns_per_frame - 1);
    if (!b) return 0.0;
    for (size_t k = 0; k < n_pk; ++k) orc_batcher_batch(b, packets + k * psz, psz, 10 + k, fr);
    orc_batcher_destroy(b);
    static const char* const names[2] = {"RANGE", "RANGE2"};
    for (int r = 0; r < 2; ++r) {
        orc_frame_field* ff = orc_frame_field_by_name(fr, names[r]);
        if (!ff) continue;
        const uint32_t* img = (const uint32_t*)ff->data;
        uint32_t* rd = destagger_new(img, shifts, fr->h, fr->w);
        void* pts = cartesian_new(img, dir, off, fr->h * fr->w, f64, omp);
        if (rd && pts)
            acc += (double)rd[fr->h * fr->w / 2] +
                   (f64 ? ((double*)pts)[fr->h * fr->w] : (double)((float*)pts)[fr->h * fr->w]);
        free(rd);
        free(pts);
    }
    return acc;
}

/* K2 baseline: packets = [F][n_pk][psz] wire bytes of F complete frames; returns seconds for `reps`
 * passes.  Each worker thread keeps one LidarFrame (as a per-sensor loop does). */
double orc_bench_k2(int mode, int f64, const orc_packet_format* pf, const uint8_t* packets, size_t F,
                    size_t n_pk, const int* shifts, const void* dir, const void* off, int threads,
                    int reps, double* sink) {
    if (threads <= 0) threads = omp_get_max_threads();
    const size_t psz = pf->lidar_packet_size;
    double acc = 0.0;
    const double t0 = now_s();
    if (mode == 2) {
#pragma omp parallel num_threads(threads) reduction(+ : acc)
        {
            orc_frame* fr = orc_frame_create(pf, 1);
            for (int rep = 0; rep < reps; ++rep) {
#pragma omp for schedule(dynamic, 1)
                for (ptrdiff_t f = 0; f < (ptrdiff_t)F; ++f)
                    acc += k2_frame(pf, fr, packets + (size_t)f * n_pk * psz, n_pk, psz, shifts, dir, off,
                                    f64, 0);
            }
            orc_frame_destroy(fr);
        }
    } else {
        if (mode == 1) omp_set_num_threads(threads);
        orc_frame* fr = orc_frame_create(pf, 1);
        for (int rep = 0; rep < reps; ++rep)
            for (size_t f = 0; f < F; ++f)
                acc += k2_frame(pf, fr, packets + f * n_pk * psz, n_pk, psz, shifts, dir, off, f64,
                                mode == 1);
        orc_frame_destroy(fr);
    }
    const double t = now_s() - t0;
    if (sink) *sink = acc;
    return t;
}

/* Parity helper for bench.py: cartesianT<float|double> + destagger<u32> of a whole pool
 * ([F][R][h][w]) into caller buffers, frames spread over the host cores (results are those of the
 * single-thread functions, bit for bit: every frame is computed by one thread). */
void orc_pool_k1(int f64, const uint32_t* rng, size_t F, size_t R, size_t h, size_t w, const int* shifts,
                 const void* dir, const void* off, void* xyz, uint32_t* rd) {
    const size_t n = h * w;
    const size_t es = f64 ? sizeof(double) : sizeof(float);
#pragma omp parallel for schedule(dynamic, 1)
    for (ptrdiff_t i = 0; i < (ptrdiff_t)(F * R); ++i) {
        const uint32_t* img = rng + (size_t)i * n;
        if (xyz) {
            void* o = (uint8_t*)xyz + (size_t)i * n * 3 * es;
            if (f64) orc_cartesian_f64((double*)o, img, (const double*)dir, (const double*)off, n);
            else orc_cartesian_f32((float*)o, img, (const float*)dir, (const float*)off, n);
        }
        if (rd && shifts) orc_destagger(sizeof(uint32_t), 1, img, shifts, h, h, w, 0, rd + (size_t)i * n);
    }
}
