// Stand-in (test infrastructure): src/testbed_nerf.cu only aliases the namespace.
#pragma once
namespace filesystem { class path {}; }
