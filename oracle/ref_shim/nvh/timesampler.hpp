#pragma once
namespace nvh { struct Stopwatch { double elapsed() { return 0.0; } }; }
