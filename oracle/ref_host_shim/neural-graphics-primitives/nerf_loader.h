// Stand-in (test infrastructure) for neural-graphics-primitives/nerf_loader.h: only the per-image metadata record the kernels of
// src/testbed_nerf.cu read (field names and defaults as nerf_loader.h:33-48; the loader itself is not needed on this path).
#pragma once
#include <neural-graphics-primitives/common.h>
#include <neural-graphics-primitives/common_device.cuh>
NGP_NAMESPACE_BEGIN
struct TrainingImageMetadata {
	const void* pixels = nullptr;
	EImageDataType image_data_type = EImageDataType::Half;
	const float* depth = nullptr;
	const Ray* rays = nullptr;
	CameraDistortion camera_distortion = {};
	Eigen::Vector2i resolution = Eigen::Vector2i::Zero();
	Eigen::Vector2f principal_point = Eigen::Vector2f::Constant(0.5f);
	Eigen::Vector2f focal_length = Eigen::Vector2f::Constant(1000.f);
	Eigen::Vector4f rolling_shutter = Eigen::Vector4f::Zero();
	Eigen::Vector3f light_dir = Eigen::Vector3f::Constant(0.f);
};
NGP_NAMESPACE_END
