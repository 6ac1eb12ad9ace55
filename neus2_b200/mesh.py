"""Mesh extraction glue (Testbed::marching_cubes + save_mesh): SDF lattice -> marching cubes -> vertex normals / colours, and
the OBJ / PLY writers of src/marching_cubes.cu:824-957 (vertices are written in DATASET coordinates: (v - offset) / scale,
faces with reversed winding, exactly as the reference does).  Host logic only; the kernels are in csrc/mesh.cu."""
import numpy as np


def extract(ctx, res, aabb=(0, 0, 0, 1, 1, 1), thresh=0.0, use_inference_params=True):
    """Testbed::marching_cubes (testbed_nerf.cu:4175-4226): the SDF itself is the 'density' (inside test sdf > thresh)."""
    lat = ctx.sdf_lattice(res, aabb, use_inference_params=use_inference_params)
    verts, faces = ctx.marching_cubes(lat, aabb, thresh)
    normals, colors = ctx.mesh_attributes(len(verts), use_inference_params=use_inference_params)
    return dict(V=verts, N=normals, C=colors, F=faces)


def _normalized(v):
    n = np.linalg.norm(v, axis=1, keepdims=True)
    return np.where(n > 0, v / np.maximum(n, 1e-30), v)


def save(filename, V, N, C, F, nerf_scale=1.0, nerf_offset=(0.0, 0.0, 0.0)):
    """save_mesh (marching_cubes.cu:824-957) without uv unwrapping: .ply (ascii) or anything else = .obj."""
    P = (np.asarray(V, dtype=np.float32) - np.asarray(nerf_offset, dtype=np.float32)) / np.float32(nerf_scale)
    Nn = _normalized(np.asarray(N, dtype=np.float32))
    C = np.asarray(C, dtype=np.float32)
    F = np.asarray(F, dtype=np.int64)
    with open(filename, "w") as f:
        if filename.lower().endswith(".ply"):
            f.write("ply\nformat ascii 1.0\ncomment output from neus2_b200\n"
                    f"element vertex {len(P)}\nproperty float x\nproperty float y\nproperty float z\nproperty float nx\nproperty float ny\nproperty float nz\n"
                    f"property uchar red\nproperty uchar green\nproperty uchar blue\nelement face {len(F)}\nproperty list uchar int vertex_index\nend_header\n")
            c8 = np.clip(C * 255.0, 0.0, 255.0).astype(np.uint8)
            for p, n, c in zip(P, Nn, c8):
                f.write("%0.5f %0.5f %0.5f %0.3f %0.3f %0.3f %d %d %d\n" % (p[0], p[1], p[2], n[0], n[1], n[2], c[0], c[1], c[2]))
            for t in F:
                f.write("3 %d %d %d\n" % (t[2], t[1], t[0]))
        else:
            cc = np.clip(C, 0.0, 1.0)
            for p, c in zip(P, cc):
                f.write("v %0.5f %0.5f %0.5f %0.3f %0.3f %0.3f\n" % (p[0], p[1], p[2], c[0], c[1], c[2]))
            for n in Nn:
                f.write("vn %0.5f %0.5f %0.5f\n" % (n[0], n[1], n[2]))
            for t in F:
                f.write("f %d//%d %d//%d %d//%d\n" % (t[2] + 1, t[2] + 1, t[1] + 1, t[1] + 1, t[0] + 1, t[0] + 1))


def load_obj(filename):
    """Minimal reader for the files `save` writes (tests / evaluation)."""
    V, C, N, F = [], [], [], []
    with open(filename) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "v":
                V.append([float(x) for x in t[1:4]]); C.append([float(x) for x in t[4:7]])
            elif t[0] == "vn":
                N.append([float(x) for x in t[1:4]])
            elif t[0] == "f":
                F.append([int(x.split("/")[0]) - 1 for x in t[1:4]])
    return np.array(V, np.float32), np.array(N, np.float32), np.array(C, np.float32), np.array(F, np.int64)
