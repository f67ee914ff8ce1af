// Stand-in for <yaml-cpp/yaml.h> (TEST INFRASTRUCTURE ONLY) for the matching.cpp harness (oracle/ref_src/matching_ref.cpp):
// node["key"], node.as<T>(), `if (node["key"])`, and YAML::LoadFile returning the tree the harness prepared.
#pragma once
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
namespace YAML {
class Node {
  public:
    Node() : d_(std::make_shared<Data>()) {}
    template <typename T> explicit Node(const T &scalar) : d_(std::make_shared<Data>()) { std::ostringstream o; o << scalar; d_->scalar = o.str(); d_->defined = true; }
    Node operator[](const std::string &key) const { auto it = d_->map.find(key); return it == d_->map.end() ? Node() : it->second; }
    Node operator[](const char *key) const { return (*this)[std::string(key)]; }
    Node operator[](int i) const { return d_->seq.at((size_t)i); }
    template <typename T> T as() const { std::istringstream in(d_->scalar); T v{}; in >> v; return v; }
    explicit operator bool() const { return d_->defined; }
    Node &set(const std::string &key, const Node &v) { d_->map[key] = v; d_->defined = true; return *this; }
  private:
    struct Data { std::string scalar; std::map<std::string, Node> map; std::vector<Node> seq; bool defined = false; };
    std::shared_ptr<Data> d_;
};
template <> inline std::string Node::as<std::string>() const { return d_->scalar; }
inline Node &b2_loadfile_tree() { static Node n; return n; }
inline Node LoadFile(const std::string &) { return b2_loadfile_tree(); }
}  // namespace YAML
