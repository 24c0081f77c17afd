// oracle/tools/dump_mesh.cpp -- TEST INFRASTRUCTURE ONLY. Build-container tool (needs /root/reference).
// Runs the reference's own OBJ path (obj_parse, src/render/obj.cpp:7; positions pass through float, q12) and writes
// the un-indexed triangle soup as raw little-endian float32 xyz triples: <out>.f32  (n_vertices * 3 floats).
// Indices are 0..n-1 by construction (obj.cpp:44-51) and are not stored.
#include <stdio.h>
#include <light_array.h>
#include "render/obj.h"

int main(int argc, char** argv) {
	if (argc != 3) { fprintf(stderr, "usage: dump_mesh in.obj out.f32\n"); return 2; }
	Vertex* vertices;
	u32* indices;
	obj_parse(argv[1], &vertices, &indices);
	u32 n = (u32)array_length(vertices);
	for (u32 i = 0; i < array_length(indices); ++i) {
		if (indices[i] != i) { fprintf(stderr, "unexpected index layout\n"); return 1; }
	}
	FILE* f = fopen(argv[2], "wb");
	for (u32 i = 0; i < n; ++i) {
		float p[3] = {vertices[i].position.x, vertices[i].position.y, vertices[i].position.z};
		fwrite(p, sizeof(float), 3, f);
	}
	fclose(f);
	fprintf(stderr, "%s: %u vertices, %u triangles\n", argv[1], n, n / 3);
	return 0;
}
