// oracle-build stand-in: the reference only needs the name Dune::Concept::Geometry to exist.
#pragma once
namespace Dune::Concept {
template <class G>
concept Geometry = true;
}
