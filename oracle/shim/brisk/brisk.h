// ORACLE SCAFFOLDING (test infrastructure): reduced umbrella header placed in
// front of the reference include dir, because internal/harris-scores.h pulls
// <brisk/brisk.h>, and the real one drags in cameras/* (OpenCV calib3d).
#pragma once
#include <brisk/brisk-descriptor-extractor.h>
#include <brisk/brisk-feature-detector.h>
#include <brisk/harris-score-calculator.h>
#include <agast/wrap-opencv.h>
#include <brisk/scale-space-feature-detector.h>
