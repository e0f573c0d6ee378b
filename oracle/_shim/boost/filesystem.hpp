// Stand-in for boost/filesystem.hpp -- oracle/_ref build only: the reference's g2o reader (d2pgo/test/posegraph_g2o.cpp) uses
// path / exists / is_directory / directory_iterator / is_regular_file, which std::filesystem provides under the same names.
#pragma once
#include <filesystem>
namespace boost { namespace filesystem = std::filesystem; }
