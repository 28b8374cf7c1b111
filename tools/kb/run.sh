#!/bin/bash
cd tools/kb
./pipe_bench_s4_t256 1080 1920 100 2
./pipe_bench_s3_t256 1080 1920 100 2
./pipe_bench_s4_t128 1080 1920 100 4
./pipe_bench_s6_t128 1080 1920 100 3
./pipe_bench_s3_t128 1080 1920 100 5
./pipe_bench_s4_t256 2160 3840 50 2
./pipe_bench_s4_t128 2160 3840 50 4
./pipe_bench_s3_t128 2160 3840 50 5
