// TEST INFRASTRUCTURE ONLY (oracle build). gfx/renderer.h names SDL_Window only as an opaque pointer type.
#pragma once
struct SDL_Window;
