// gltf_import.h — glTF 2.0 -> flat scene arrays in the shape nvh::GltfScene leaves behind after
// importMaterials + importDrawableNodes (reference scene.cpp:60-76).  Replaces tinygltf + nvh::GltfScene
// (both un-vendored third-party code of the reference).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "eidola.h"

namespace eid {

struct HostGltf {
  std::vector<float> positions, normals, tangents, texcoords0, colors0;   // 3,3,4,2,4 floats per vertex
  std::vector<uint32_t> indices;
  std::vector<eid_prim_mesh> primMeshes;
  std::vector<eid_node> nodes;
  std::vector<eid_material_desc> materials;
  std::vector<eid_light_desc> lights;
  struct Image { uint32_t width = 0, height = 0; std::vector<uint8_t> rgba8; };   // decoded, empty = not available
  std::vector<Image> images;
  std::vector<eid_texture_desc> textures;
  bool hasCamera = false;
  float camEye[3] = {0, 0, 0}, camCenter[3] = {0, 0, -1}, camUp[3] = {0, 1, 0};
  float camYfovRad = 1.0471975512f;
  float bboxMin[3] = {0, 0, 0}, bboxMax[3] = {0, 0, 0};   // m_dimensions (world space)

  void fromDesc(const eid_scene_desc& d);
  void computeDimensions();
};

// PNG -> RGBA8 with stb_image's conventions (png_decode.cpp); throws eid::Error(EID_ERR_PARSE) on a malformed file
bool isPng(const uint8_t* data, size_t bytes);
void decodePng(const uint8_t* data, size_t bytes, HostGltf::Image& out);

// throws eid::Error (EID_ERR_IO / EID_ERR_PARSE / EID_ERR_UNSUPPORTED)
// `provided` = images the host decoded beforehand (index -> texels).  PNG images (files, data URIs, .glb bufferViews) are decoded by the
// library; a file whose used textures reference any other undecoded image (JPEG ...) is rejected with EID_ERR_UNSUPPORTED.
void importGltfFile(const std::string& path, HostGltf& out, const std::vector<HostGltf::Image>& provided);

}  // namespace eid
