// Sanitizer harness for the headless host's file parsers (PNG, JPEG, glTF / GLB, OBJ, TOML): decodes every file named on the command
// line, counting refusals; built with -fsanitize=address,undefined by tests/test_host_fuzz.py.
#include <cstdio>
#include <fstream>
#include <string>
#include <vector>
#include "config.h"
#include "gltf.h"
#include "texture.h"
using namespace rbhost;
int main(int argc, char** argv) {
    int ok = 0, bad = 0;
    for (int i = 1; i < argc; i++) {
        std::string p = argv[i];
        try {
            if (p.size() > 4 && (p.substr(p.size() - 4) == ".glb" || p.substr(p.size() - 5) == ".gltf")) { Scene s = load_gltf_scene(p); SceneTables t = s.build(false); (void)t; }
            else if (p.size() > 4 && p.substr(p.size() - 4) == ".obj") { ModelData md = load_obj(p); Scene s; s.addObject(md, identity(), Material()); SceneTables t = s.build(false); (void)t; }
            else if (p.size() > 5 && p.substr(p.size() - 5) == ".toml") { Config c = Config::from_toml(TomlDoc::parse_file(p)); (void)c; }
            else { Image8 im = load_image_rgba8(p, true); (void)im; }
            ok++;
        } catch (const std::exception& e) { bad++; }
    }
    std::printf("ok %d refused %d\n", ok, bad);
    return 0;
}
