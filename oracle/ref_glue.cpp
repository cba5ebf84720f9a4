// TEST INFRASTRUCTURE ONLY.  Host-side glue for oracle/_ref: the reference defines the tinyobj
// implementation inside main.cpp (main.cpp:4-5), which cannot be compiled here (Windows.h, GL,
// OpenCV, libtorch), and declares Scene::~Scene (scene.h:21) without ever defining it.
#define TINYOBJLOADER_IMPLEMENTATION
#include "tiny_obj_loader.h"
#include "scene.h"
Scene::~Scene() {}
