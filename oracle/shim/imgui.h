// TEST INFRASTRUCTURE ONLY (oracle build). The few Dear ImGui types gfx/renderer.h and gfx/waveform_visual.cpp
// name in declarations; nothing here is ImGui code and none of it is executed by the oracle.
#pragma once
struct ImVec2 { float x = 0, y = 0; ImVec2() = default; ImVec2(float a, float b) : x(a), y(b) {} };
struct ImVec4 { float x = 0, y = 0, z = 0, w = 0; ImVec4() = default; ImVec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {} };
struct ImGuiViewport { void* RendererUserData = nullptr; };
struct ImDrawData;
struct ImDrawList;
struct ImDrawCmd;
typedef unsigned int ImU32;
typedef unsigned short ImDrawIdx;
typedef void* ImTextureID;
