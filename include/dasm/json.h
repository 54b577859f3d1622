// Minimal JSON reader with the boost::property_tree-like surface the reference's factories use
// (params.get<T>(key, default), try_get_child; include/json.h:6-17 of the reference; boost is not available here).
// Values are kept as strings (the reference's JSON files store numbers as strings or numbers interchangeably).
#pragma once
#include <cctype>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace dasm
{
  class ptree
  {
  public:
    std::string                                  value;
    std::vector<std::pair<std::string, ptree>>   children;

    const ptree *
    find(const std::string &key) const
    {
      for (const auto &c : children)
        if (c.first == key)
          return &c.second;
      return nullptr;
    }

    template <typename T>
    T
    get(const std::string &key, const T &def) const
    {
      const ptree *c = find(key);
      return c ? convert<T>(c->value) : def;
    }

    std::string
    get(const std::string &key, const char *def) const
    {
      const ptree *c = find(key);
      return c ? c->value : std::string(def);
    }

    template <typename T>
    T
    get(const std::string &key) const
    {
      const ptree *c = find(key);
      if (!c)
        throw std::runtime_error("No such node (" + key + ")");
      return convert<T>(c->value);
    }

    template <typename T>
    void
    put(const std::string &key, const T &v)
    {
      std::ostringstream ss;
      ss << std::boolalpha << v;
      for (auto &c : children)
        if (c.first == key)
          {
            c.second.value = ss.str();
            return;
          }
      ptree t;
      t.value = ss.str();
      children.emplace_back(key, t);
    }

    void
    add_child(const std::string &key, const ptree &child)
    {
      children.emplace_back(key, child);
    }

    static ptree
    parse(const std::string &text)
    {
      size_t pos = 0;
      ptree  t   = parse_value(text, pos);
      return t;
    }

    static ptree
    parse_file(const std::string &file)
    {
      std::ifstream f(file);
      if (!f)
        throw std::runtime_error("cannot open " + file);
      std::stringstream ss;
      ss << f.rdbuf();
      return parse(ss.str());
    }

  private:
    template <typename T>
    static T
    convert(const std::string &s)
    {
      if constexpr (std::is_same<T, bool>::value)
        return s == "true" || s == "1";
      else if constexpr (std::is_same<T, std::string>::value)
        return s;
      else
        {
          std::istringstream ss(s);
          T                  v{};
          ss >> v;
          if (ss.fail())
            throw std::runtime_error("conversion of <" + s + "> failed");
          return v;
        }
    }

    static void
    skip(const std::string &t, size_t &p)
    {
      while (p < t.size() && std::isspace((unsigned char)t[p]))
        ++p;
    }

    static std::string
    parse_string(const std::string &t, size_t &p)
    {
      std::string out;
      ++p; // opening quote
      while (p < t.size() && t[p] != '"')
        {
          if (t[p] == '\\' && p + 1 < t.size())
            ++p;
          out += t[p++];
        }
      ++p;
      return out;
    }

    static ptree
    parse_value(const std::string &t, size_t &p)
    {
      skip(t, p);
      ptree node;
      if (p >= t.size())
        throw std::runtime_error("JSON: unexpected end");
      if (t[p] == '{')
        {
          ++p;
          skip(t, p);
          while (p < t.size() && t[p] != '}')
            {
              skip(t, p);
              const std::string key = parse_string(t, p);
              skip(t, p);
              if (t[p] != ':')
                throw std::runtime_error("JSON: expected ':'");
              ++p;
              node.children.emplace_back(key, parse_value(t, p));
              skip(t, p);
              if (t[p] == ',')
                ++p;
              skip(t, p);
            }
          ++p;
        }
      else if (t[p] == '[')
        {
          ++p;
          skip(t, p);
          while (p < t.size() && t[p] != ']')
            {
              node.children.emplace_back("", parse_value(t, p));
              skip(t, p);
              if (t[p] == ',')
                ++p;
              skip(t, p);
            }
          ++p;
        }
      else if (t[p] == '"')
        node.value = parse_string(t, p);
      else
        {
          while (p < t.size() && t[p] != ',' && t[p] != '}' && t[p] != ']' && !std::isspace((unsigned char)t[p]))
            node.value += t[p++];
        }
      return node;
    }
  };

  // include/json.h:6-17 of the reference
  inline ptree
  try_get_child(const ptree &params, const std::string &label)
  {
    const ptree *c = params.find(label);
    return c ? *c : ptree();
  }
} // namespace dasm
