#!/usr/bin/env python3
"""Informative CPU line of bench.py (never the reference arm): what REAL OpenCV (cv2: SIMD / IPP dispatch) needs for the dense
stages of ORBextractor::operator() alone -- the 8-level pyramid (resize + copyMakeBorder), FAST per level, GaussianBlur per
level -- one process per core, cv2 single-threaded inside each.  Leaves out the quadtree, orientation, descriptors and stereo
matching, and runs FAST once per level at iniThFAST instead of per 30-px cell with the minThFAST retry.
Usage: cv2_dense_baseline.py FRAMES.npy CORES LEVELS SCALE INI_TH   -> one JSON line {"frames_per_s": ..., "ms_per_frame_per_core": {...}}"""
import json
import multiprocessing as mp
import sys
import time

import numpy as np


def work(args):
    import cv2
    path, lo, hi, levels, scale, ini_th = args
    cv2.setNumThreads(1)
    frames = np.load(path, mmap_mode="r")
    inv = [np.float32(1.0)]
    for _ in range(1, levels):
        inv.append(np.float32(inv[-1] / np.float32(scale)))
    det = cv2.FastFeatureDetector_create(threshold=ini_th, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    T = dict(pyramid=0.0, fast=0.0, blur=0.0)
    for i in range(lo, hi):
        img = np.ascontiguousarray(frames[i])
        h, w = img.shape
        lvl = img
        for l in range(levels):
            t = time.perf_counter()
            if l:
                lvl = cv2.resize(lvl, (int(round(w * float(inv[l]))), int(round(h * float(inv[l])))), interpolation=cv2.INTER_LINEAR)
            cv2.copyMakeBorder(lvl, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
            T["pyramid"] += time.perf_counter() - t
            t = time.perf_counter()
            det.detect(lvl)
            T["fast"] += time.perf_counter() - t
            t = time.perf_counter()
            cv2.GaussianBlur(lvl, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
            T["blur"] += time.perf_counter() - t
    return T, hi - lo


if __name__ == "__main__":
    path, cores, levels, scale, ini_th = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5])
    n = np.load(path, mmap_mode="r").shape[0]
    cuts = [n * k // cores for k in range(cores + 1)]
    jobs = [(path, cuts[k], cuts[k + 1], levels, scale, ini_th) for k in range(cores) if cuts[k + 1] > cuts[k]]
    with mp.Pool(len(jobs)) as pool:
        pool.map(work, [(path, 0, 1, levels, scale, ini_th)] * len(jobs))          # warm-up: import cv2, first calls
        t0 = time.perf_counter()
        res = pool.map(work, jobs)
        dt = time.perf_counter() - t0
    tot = {k: sum(r[0][k] for r in res) for k in res[0][0]}
    print(json.dumps({"frames_per_s": n / dt, "ms_per_frame_per_core": {k: 1e3 * v / n for k, v in tot.items()}}))
